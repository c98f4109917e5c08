/*
 * cae_b200.h - C ABI of libcae_b200.so: the sm_100a kernels behind the cae_tools hot path.
 *
 * The reference (surftemp/cae_tools) has no FFI of its own: its hot path is a chain of
 * PyTorch ATen calls made from nn.Module.forward / autograd / torch.optim.  Each entry
 * point below names the reference call site (path:line under the reference tree) whose
 * arithmetic it replaces.  Plain pointers, ints and PODs only - loadable with ctypes/cffi.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CAE_E* code for bad arguments or a
 *     positive cudaError_t; cae_last_error() returns a thread-local message.
 *   - all pointers are DEVICE pointers unless stated; the library never allocates
 *     persistent device memory - callers (PyTorch's allocator) own every buffer.
 *   - kernels are enqueued on the stream given (a cudaStream_t passed as void*); no call
 *     synchronises, so everything is CUDA-graph capturable.
 *   - tensors are fp32 NCHW with explicit strides (CaeView) so that internal buffers may
 *     use padded rows and channel-offset views (skip concatenation writes in place).
 */
#ifndef CAE_B200_H
#define CAE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CAE_OK            0
#define CAE_EINVAL       -1   /* bad argument */
#define CAE_EUNSUPPORTED -2   /* geometry outside what the kernels implement */

/* 4-D fp32 view: element (n,c,y,x) lives at p[n*sN + c*sC + y*ld + x] */
typedef struct CaeView {
    float*    p;
    int       N, C, H, W;
    int       ld;
    long long sC, sN;
} CaeView;

/* Operand read through an on-load transform:
 *     v = k0[c]*t0 + k1[c]*t1 + k2[c] ;  if (relu) v = max(v,0)
 * NULL k0/k1/k2 mean 1/0/0; t1 (same geometry as t0) may be NULL.  Used to apply
 * BatchNorm+ReLU of the producing layer while loading (forward) and to form
 * dL/dy = A*dz + B*y + C of a BatchNorm backward while loading (backward), so neither
 * is ever a standalone pass over HBM.
 * cursor (device int*, may be NULL) selects the current batch of a pre-batched
 * device-resident data set: the base address is advanced by cursor[0]*cursor_stride
 * elements (reference keeps all batches on the device: conv_ae_model.py:315-325). */
typedef struct CaeSrc {
    CaeView       t0;
    const float*  t1;
    const float*  k0;
    const float*  k1;
    const float*  k2;
    int           relu;
    const int*    cursor;
    long long     cursor_stride;
    const float*  kn;             /* optional per-(sample, channel) multiplier of the t0 term, [N][C] (channel-attention
                                     gate, unet.py:158-159); honoured by cae_ew_epilogue only */
} CaeSrc;

typedef struct CaeConvGeom {
    int kh, kw, stride, pad;
} CaeConvGeom;

/* Per-layer BatchNorm state; every pointer addresses C floats unless noted.
 * (reference: nn.BatchNorm2d in encoder.py:45, decoder.py:47; eps 1e-5, momentum 0.1) */
typedef struct CaeBN {
    int          C;
    float        eps, momentum;
    const float* gamma;
    const float* beta;
    float*       running_mean;
    float*       running_var;
    long long*   num_batches_tracked;   /* 1 x int64 */
    float*       scale;     /* y_hat = scale*y + shift   (written by fwd finalize / eval prepare) */
    float*       shift;
    float*       mean;      /* batch mean / 1/sqrt(var+eps) of the last training forward */
    float*       invstd;
    float*       dgamma;    /* gradients, written by bwd finalize (may point into the flat grad arena) */
    float*       dbeta;
    float*       dbias;     /* grad of the bias of the conv feeding this BN (may be NULL) */
    float*       bwdA;      /* dL/dy = A*dz + B*y + C  (written by bwd finalize) */
    float*       bwdB;
    float*       bwdC;
} CaeBN;

/* Epilogue of the conv kernels (what happens to each accumulated output element) */
#define CAE_EPI_PLAIN      0  /* out = acc + bias                                                  */
#define CAE_EPI_STATS      1  /* out = acc + bias ; per-channel sum/sumsq -> BatchNorm (training)   */
#define CAE_EPI_MASKSTATS  2  /* out = acc * [relu mask of `act`] ; sums for BatchNorm backward     */
#define CAE_EPI_SIGMOID    3  /* out = sigmoid(acc + bias)                                          */
#define CAE_EPI_SIGMOID_MSE 4 /* yhat = sigmoid(acc+bias); loss += (yhat-t)^2 ; out = dL/d(acc)     */
#define CAE_EPI_MASK       5  /* out = acc * [act > 0]  (ReLU backward without a BatchNorm)         */

typedef struct CaeEpilogue {
    int            mode;
    const float*   bias;          /* per output channel, may be NULL */
    /* STATS / MASKSTATS / SIGMOID_MSE: deterministic two-stage reduction */
    double*        partials;      /* workspace >= grid_x * C * 2 doubles (see cae_partials_len) */
    unsigned int*  ticket;        /* 1 x u32, zero before first use; self-resetting */
    CaeBN          bn;            /* STATS: BN that follows this conv. MASKSTATS: BN whose backward is reduced */
    CaeView        act;           /* MASKSTATS: pre-BN output y of the layer whose ReLU/BN is differentiated
                                     (same geometry as the kernel output). act.p NULL = no mask, plain write */
    /* SIGMOID_MSE */
    CaeSrc         target;        /* same geometry as the kernel output */
    float*         loss_out;      /* loss_out[cursor ? *cursor : 0] = mean squared error of this batch */
    float*         dbias;         /* gradient of `bias` (C floats) */
    int            write_mode;    /* 0: write dL/d(acc) ; 1: write yhat ; 2: write nothing (loss only) */
    float          count_scale;   /* loss and dL/d(acc) are multiplied by this (0 means 1): N_local/N_global when the
                                     batch is sharded over ranks, so that SUM-all-reduced gradients and losses are
                                     those of the global batch */
    CaeSrc         addend;        /* optional (addend.t0.p != NULL): a second gradient of the output's geometry added to acc
                                     before the epilogue acts - the skip-connection fan-in of the UNET encoder */
} CaeEpilogue;

const char* cae_last_error(void);
/* 1: generic direct kernels only (v1); 2 (default): tiled shared-memory kernels where they apply */
void cae_set_kernel_generation(int gen);
int  cae_version(void);
/* sizeof() of the PODs of this header as the library was compiled (binding self-check): which = 0 CaeView, 1 CaeSrc,
 * 2 CaeConvGeom, 3 CaeBN, 4 CaeEpilogue, 5 CaeGemm, 6 CaePatchHead, 7 CaeFcStack, 8 CaeUnetStem; -1 otherwise */
long long cae_struct_size(int which);
/* number of doubles the `partials` workspace must hold for a kernel whose output has C channels */
long long cae_partials_len(int C);

/* ---- convolution family -------------------------------------------------------------
 * cae_conv_down: strided convolution, weights [Cout][Cin][kh][kw]
 *     out[n,co,oy,ox] = sum_{ci,ky,kx} in(n,ci,oy*s+ky-p,ox*s+kx-p) * W[co,ci,ky,kx]
 *   forward of nn.Conv2d (reference encoder.py:43-44, unet.py:81-82) and, with the same weight
 *   tensor, the input-gradient of nn.ConvTranspose2d (autograd of decoder.py:44-45).
 * cae_conv_up: transposed convolution (gather form), weights [Cin][Cout][kh][kw]
 *     out[n,co,oy,ox] = sum_{ci} sum_{ky,kx : (oy+p-ky)%s==0 ...} in(n,ci,(oy+p-ky)/s,(ox+p-kx)/s) * W[ci,co,ky,kx]
 *   forward of nn.ConvTranspose2d (decoder.py:44-45, unet.py:138-140; followed by torch.sigmoid
 *   decoder.py:77 and nn.MSELoss conv_ae_model.py:193 through the epilogue) and the
 *   input-gradient of nn.Conv2d.
 * cae_conv_wgrad: weight gradient of either, G[cs][cb][ky][kx] =
 *     sum_{n,i,j} small(n,cs,i,j) * big(n,cb,i*s+ky-p,j*s+kx-p)
 *   (convT: small = layer input, big = dL/dy; conv: small = dL/dy, big = layer input).
 *   `partials` needs cae_wgrad_partials_len() floats. */
int cae_conv_down(const CaeSrc* in, const float* weight, const CaeConvGeom* g, const CaeView* out,
                  const CaeEpilogue* epi, void* stream);
int cae_conv_up(const CaeSrc* in, const float* weight, const CaeConvGeom* g, const CaeView* out,
                const CaeEpilogue* epi, void* stream);
int cae_conv_wgrad(const CaeSrc* small_op, const CaeSrc* big_op, const CaeConvGeom* g, float* grad,
                   float* partials, unsigned int* ticket, void* stream);
long long cae_wgrad_partials_len(const CaeSrc* small_op, const CaeSrc* big_op, const CaeConvGeom* g);

/* Elementwise member of the family: out = epilogue(in) - used where a gradient arrives from a
 * non-conv producer (the fc stack) and still needs the ReLU mask + BatchNorm-backward sums. */
int cae_ew_epilogue(const CaeSrc* in, const CaeView* out, const CaeEpilogue* epi, void* stream);

/* ---- fully connected ------------------------------------------------------------------
 * C[m,n] = epi( sum_k A(m,k) * B(k,n) ), A(m,k) = Ap[m*sAm + k*sAk], B(k,n) = Bp[k*sBk + n*sBn]
 * replaces aten::addmm / mm of nn.Linear forward and backward (encoder.py:54-58, decoder.py:31-35).
 *   a_k0/a_k2 (may be NULL): A is read as relu?(a_k0[k/a_hw]*A + a_k2[k/a_hw])  (BatchNorm+ReLU+Flatten on load)
 *   b_k0/b_k2: same for B with channel = k... see cae_gemm fields.
 *   bias (per n, may be NULL), relu_out, mask (same layout as C: C *= mask>0), bias_grad (per m: sum_k A(m,k))
 */
typedef struct CaeGemm {
    int M, N, K;
    const float* A; long long sAm, sAk;
    const float* B; long long sBk, sBn;
    float*       C; long long sCm, sCn;
    /* on-load per-channel affine (+relu) for A indexed by k, or for B indexed by n */
    const float* a_k0; const float* a_k2; int a_hw; int a_relu;   /* channel = k / a_hw */
    const float* b_k0; const float* b_k2; int b_hw; int b_relu;   /* channel = n / b_hw */
    const float* bias;        /* + bias[n] */
    int          relu_out;    /* C = max(C,0) */
    const float* mask;        /* C *= (mask[m*sCm+n*sCn] > 0) */
    float*       rowsum_A;    /* if non-NULL: rowsum_A[m] = sum_k A(m,k)  (bias gradient when A = dy^T) */
} CaeGemm;
int cae_gemm(const CaeGemm* g, void* stream);
/* The same contraction on the tensor cores (tc_dense.cu: split -> cae_tc_gemm 3xTF32 -> epilogue), for the nn.Linear layers
 * of the large-fc regimes (unet.py:92-100,121-129 at fc 3200 / latent 800 and batch 256; linear.py:43 at large batches).
 * cae_gemm_tc_workspace: floats of caller-owned scratch the call needs, 0 when the problem is not eligible (min(M,N) < 128,
 * K < 32, < 0.1 GFLOP, an operand contiguous along neither axis, C not row-major): the caller then uses cae_gemm. */
long long cae_gemm_tc_workspace(const CaeGemm* g);
int cae_gemm_tc(const CaeGemm* g, float* workspace, long long workspace_len, void* stream);

/* ---- batch norm helpers ------------------------------------------------------------------ */
/* eval mode: scale/shift from running statistics for `count` layers (table lives on the DEVICE) */
int cae_bn_eval_prepare(const CaeBN* device_table, int count, void* stream);

/* ---- loss (eval: no sigmoid fusion needed when yhat already exists) ---------------------- */
/* loss_out[slot] = mean((a-b)^2) over n elements; deterministic two-stage reduction */
int cae_mse(const float* a, const float* b, long long n, double* partials, unsigned int* ticket,
            float* loss_out, const int* cursor, void* stream);

/* ---- optimiser -----------------------------------------------------------------------------
 * fused multi-tensor Adam/AdamW over one flat arena (reference: torch.optim.Adam
 * conv_ae_model.py:310 (coupled L2) and torch.optim.AdamW unet.py:457 (decoupled)).
 * step_count: device int holding the number of steps already taken (t = *step_count + 1).
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce). */
int cae_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
             float eps, float weight_decay, int decoupled, float grad_scale, const int* step_count, void* stream);
/* end-of-step bookkeeping: step_count[0] += 1 ; if cursor: cursor[0] = (cursor[0]+1) % n_batches */
int cae_step_advance(int* step_count, int* cursor, int n_batches, void* stream);
/* cae_adam + cae_step_advance in one launch (the last CTA to finish - `ticket`, zero-initialised - does the bookkeeping) */
int cae_adam_advance(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int decoupled, float grad_scale, int* step_count, int* cursor,
                     int n_batches, unsigned int* ticket, void* stream);

/* ---- UNET pieces (reference: src/cae_tools/models/unet.py) -------------------------------------------
 * plane = one (n, c) image.  cae_plane_stats: stats[(n*C+c)*4 + {0..3}] = sum, sum of squares, max, argmax
 *   (AdaptiveAvgPool2d / AdaptiveMaxPool2d of ChannelAttention, unet.py:26-27, and the BatchNorm sums of the gate)
 * cae_channel_attention_fwd: att = sigmoid(W2 relu(W1 avg) + W2 relu(W1 max))  (unet.py:35-39); hid keeps the
 *   two hidden vectors per sample for the backward pass ([N][2][Cr])
 * cae_channel_attention_bwd: gradients of W1 [Cr][C], W2 [C][Cr] and of avg (per pixel) / max per plane
 * cae_plane_dot: out[n,c] = sum_hw g * y ; cae_gate_bwd: dy = att*g + davg + dmax*[argmax], plane sums of dy
 * cae_sum_over_n: out[c] = sum_n in[n][c] */
int cae_plane_stats(const CaeView* y, float* stats, void* stream);
int cae_channel_attention_fwd(const float* stats, const float* W1, const float* W2, int N, int C, int Cr, int HW,
                              float* att, float* hid, void* stream);
int cae_channel_attention_bwd(const float* datt, const float* att, const float* hid, const float* stats,
                              const float* W1, const float* W2, int N, int C, int Cr, int HW, float* dW1, float* dW2,
                              float* davg, float* dmax, void* stream);
int cae_plane_dot(const CaeSrc* g, const CaeView* y, float* out, void* stream);
int cae_gate_bwd(const CaeSrc* g, const float* att, const float* davg, const float* dmax, const float* stats,
                 const CaeView* dy, float* plane_sum, void* stream);
int cae_sum_over_n(const float* in, int N, int C, float* out, void* stream);
/* masked MSE + lambda * (1 - mean Pearson) (unet.py:314-320,635-678): pred = sigmoid output; mask may be absent
 * (mask->t0.p NULL = ones) and has 1 or C channels.  Writes loss_out[slot] = masked MSE, pearson_out[slot] =
 * 1 - mean corr; with dz != NULL also dL/d(pre-sigmoid) and its plane sums (bias gradient pieces).
 * moments: N*C*7 doubles, coef: N*C*3 floats, scalars: 3 floats of workspace.
 * mse_scale (device, may be NULL): per-batch factor of the masked-MSE term, read at the target's cursor slot, used INSTEAD of
 * count_scale for that term - data parallelism with masks: valid pixels of this rank's share / valid pixels of the
 * global batch, so that the SUM over ranks is sum((d-t)^2 m^2) / sum(m) of the global batch exactly (the Pearson term is a
 * mean over samples and keeps count_scale). */
int cae_masked_pearson_loss(const CaeView* pred, const CaeSrc* target, const CaeSrc* mask, int mask_channels,
                            float lambda_pearson, float count_scale, double* moments, float* coef, float* scalars,
                            float* loss_out, float* pearson_out, const CaeView* dz, float* plane_sum,
                            const float* mse_scale, void* stream);

/* ---- fc bottleneck in one launch per direction (replaces the cae_gemm / cae_ew_epilogue chain when it fits one CTA) --
 * ConvAEModel (encoder.py:52-58, decoder.py:29-35):  Linear ReLU Linear | Linear ReLU Linear           (bn*.C == 0, relu_mid 0)
 * UNET (unet.py:92-100,121-129): Linear BatchNorm1d ReLU Linear ReLU | Linear BatchNorm1d ReLU Linear ReLU   (relu_mid 1)
 * A [N][in1] is read through an optional per-channel affine + ReLU (BatchNorm2d + ReLU + Flatten of the last encoder
 * layer: channel = k / a_hw).  Saved for the backward pass: t1 / t3 (pre-activations of Linear 1 / 3), z, u.
 * Backward: du [N][out4] must already carry the mask of the last activation; writes every weight / bias gradient,
 * the BatchNorm gradients (bn*.dgamma / dbeta) and dA [N][in1]; biases that feed a BatchNorm get no gradient written
 * (identically zero). */
typedef struct CaeFcStack {
    int          N, in1, fc1, lat, fc2, out4;
    const float* A; const float* a_k0; const float* a_k2; int a_hw; int a_relu;
    const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4;       /* nn.Linear layout [out][in] */
    CaeBN        bn1, bn3;
    int          train;        /* BatchNorm: batch statistics (1) or the eval coefficients already in bn.scale / shift (0) */
    int          relu_mid;
    float        *t1, *z, *t3, *u;
    const float* du;
    float        *dW1, *db1, *dW2, *db2, *dW3, *db3, *dW4, *db4, *dA;
} CaeFcStack;
int cae_fc_stack_supported(int N, int in1, int fc1, int lat, int fc2, int out4);
int cae_fc_stack_fwd(const CaeFcStack* p, void* stream);
int cae_fc_stack_bwd(const CaeFcStack* p, void* stream);

/* ---- fused attention block of the UNET decoder (unet.py:23-39,149-163): between two transposed convolutions
 *   forward : plane statistics of y (AdaptiveAvg/MaxPool), att = sigmoid(W2 relu(W1 avg) + W2 relu(W1 max)),
 *             cat = [att*y ; skip] and - epilogue STATS - the BatchNorm2d(2C) statistics of cat, in ONE launch
 *             (replaces cae_plane_stats + cae_channel_attention_fwd + two cae_ew_epilogue launches)
 *   backward: dL/d att, the MLP backward (dW1, dW2), dL/dy = att*g + davg + dmax*[argmax] and dbias = sum dL/dy in ONE
 *             launch (replaces cae_plane_dot + cae_channel_attention_bwd + cae_gate_bwd + cae_sum_over_n)
 * One CTA per sample with the sample's planes in shared memory: cae_attention_block_supported() says whether a
 * geometry fits; larger ones use the unfused entry points above.  stats [N*C*4], att [N*C], hid [N*2*Cr] are saved for
 * the backward call; partials: cae_attention_block_partials_len() floats. */
int       cae_attention_block_supported(int C, int H, int W, int Cr);
long long cae_attention_block_partials_len(int C, int Cr);
int       cae_attention_block_fwd(const CaeView* y, const CaeSrc* skip, const float* W1, const float* W2, int Cr,
                                  const CaeView* cat, const CaeEpilogue* epi, float* stats, float* att, float* hid,
                                  void* stream);
int       cae_attention_block_bwd(const CaeSrc* g, const CaeView* y, const float* att, const float* hid,
                                  const float* stats, const float* W1, const float* W2, int Cr, const CaeView* dy,
                                  float* dW1, float* dW2, float* dbias, float* partials, unsigned int* ticket,
                                  void* stream);

/* ---- eval-mode UNET stem: every layer before the last transposed convolution in ONE launch (BatchNorm in eval mode
 * is a per-channel affine, so samples are independent: a CTA carries a few samples through the whole stem in shared
 * memory).  Used by apply() / score() / the test epoch.  scale / shift are the eval-mode BatchNorm coefficients
 * written by cae_bn_eval_prepare (NULL = no BatchNorm).  Reference: unet.py:73-163 in eval mode.
 *   conv[l] : Conv2d(k, stride, pad) + BN + ReLU                                   weights [Cout][Cin][k][k]
 *   fc[l]   : y = act((W x + b) * scale + shift)                                   weights [out][in]
 *   up[j]   : ConvTranspose2d(k, stride, pad), ChannelAttention (W1 [Cr][C], W2 [C][Cr]), concat with the activated
 *             output of encoder layer `skip`, BN(2C) + ReLU                        weights [Cin][Cout][k][k]
 * out: [N, 2*C_last, H_last, W_last] - the ACTIVATED input of the last layer. */
#define CAE_STEM_MAX 4
typedef struct CaeStemConv {
    int Cin, Hin, Win, Cout, Hout, Wout, k, stride, pad;
    const float *w, *b, *scale, *shift;
} CaeStemConv;
typedef struct CaeStemFc {
    int in, out, relu;
    const float *w, *b, *scale, *shift;
} CaeStemFc;
typedef struct CaeStemUp {
    int Cin, Hin, Win, Cout, Hout, Wout, k, stride, pad, Cr, skip;
    const float *w, *b, *W1, *W2, *scale, *shift;
} CaeStemUp;
typedef struct CaeUnetStem {
    int n_conv, n_fc, n_up;
    CaeStemConv conv[CAE_STEM_MAX];
    CaeStemFc   fc[CAE_STEM_MAX];
    CaeStemUp   up[CAE_STEM_MAX];
} CaeUnetStem;
int cae_unet_stem_supported(const CaeUnetStem* s);
int cae_unet_stem_eval(const CaeUnetStem* s, const CaeSrc* x, const CaeView* out, void* stream);

/* ---- training-mode UNET stem in TWO launches (unet_stem_train.cu) ------------------------------------------------------
 * Everything before the last transposed convolution of UNET.__train_epoch (unet.py:295-337 over Encoder.forward :102-112
 * and Decoder.forward :149-163) - Conv2d/BatchNorm2d/ReLU/Dropout blocks, the two Linear-BatchNorm1d-ReLU-Dropout-Linear-
 * ReLU-Dropout stacks, ConvTranspose2d + ChannelAttention gate + skip concat + BatchNorm2d(2C)/ReLU/Dropout blocks - as one
 * forward and one backward COOPERATIVE kernel: a CTA owns 1..4 samples and keeps their activations (forward) and gradients
 * (backward) in shared memory; training-mode BatchNorm couples the samples, so every BatchNorm layer is one fixed-order
 * cross-CTA reduction (per-CTA partial rows in global memory + a grid barrier) - 7 barriers forward, 8 backward for the
 * shipped spec, instead of ~55 dependent launches.  The weight gradients are per-CTA partial rows summed in row order.
 * Dropout (p > 0): mask = hash(seed, step_count[0], site, sample, element) >= p, recomputed (never stored) in backward;
 * skip connections read the activation BEFORE dropout, as the reference does (in-place ReLU, unet.py:84-85,108-109).
 *   forward : writes `tape` ([N][tape_elems] raw layer outputs), the BatchNorm scale/shift/mean/invstd + running
 *             statistics of every CaeBN, and `hin` = the ACTIVATED input of the head [N, 2C, H, W]
 *   backward: reads `tape` and `dhin` (= dL/d hin, PLAIN epilogue of cae_patch_head_bwd), writes every gradient
 *             (dw / db / dW1 / dW2, dgamma / dbeta of every CaeBN; zero for the dead biases in front of a BatchNorm) */
typedef struct CaeStemTrainConv {
    int Cin, Hin, Win, Cout, Hout, Wout, k, stride, pad;
    const float *w, *b;
    float *dw, *db;
    CaeBN bn;
} CaeStemTrainConv;
typedef struct CaeStemTrainFc {
    int in, out, has_bn;       /* has_bn: Linear - BatchNorm1d - ReLU - Dropout; else Linear - ReLU - Dropout */
    const float *w, *b;
    float *dw, *db;
    CaeBN bn;
} CaeStemTrainFc;
typedef struct CaeStemTrainUp {
    int Cin, Hin, Win, Cout, Hout, Wout, k, stride, pad, Cr, skip;
    const float *w, *b, *W1, *W2;
    float *dw, *db, *dW1, *dW2;
    CaeBN bn;                  /* BatchNorm2d(2*Cout) */
} CaeStemTrainUp;
typedef struct CaeStemTrain {
    int n_conv, n_fc, n_up, N;
    CaeStemTrainConv conv[CAE_STEM_MAX];
    CaeStemTrainFc   fc[CAE_STEM_MAX];
    CaeStemTrainUp   up[CAE_STEM_MAX];
    const float* params;       /* contiguous block (16-byte aligned, multiple of 4 floats) that holds EVERY w / b / W1 / W2 /
                                  gamma / beta above: the engine's flat parameter arena up to the head's weights; copied to
                                  shared memory by one bulk async copy per launch */
    long long params_len;
    float* tape;               /* [N][cae_unet_stem_train_tape_elems] */
    float* hin;                /* [N, 2*C_last, H_last, W_last] contiguous */
    const float* dhin;
    float dropout_p;
    unsigned long long seed;
    const int* step_count;
    double* bnpart;            /* cae_unet_stem_train_workspace(.., 0) doubles */
    float* wpart;              /* cae_unet_stem_train_workspace(.., 1) floats */
} CaeStemTrain;
/* 1 when the geometry / batch fit the fused kernels (<= 4 samples per CTA on 148 CTAs, activations in shared memory) */
int       cae_unet_stem_train_supported(const CaeStemTrain* s);
long long cae_unet_stem_train_tape_elems(const CaeStemTrain* s);
long long cae_unet_stem_train_workspace(const CaeStemTrain* s, int which);
int       cae_unet_stem_train_fwd(const CaeStemTrain* s, const CaeSrc* x, void* stream);
int       cae_unet_stem_train_bwd(const CaeStemTrain* s, const CaeSrc* x, void* stream);
/* profiling aid: 64 clock64() phase timestamps of CTA 0 from the last forward ([0..31]) and backward ([32..63]) launch
 * (slot 31 / 63 = one past the last slot used) */
int       cae_unet_stem_train_profile(unsigned long long* out64);

/* ---- patch head: transposed convolution with kernel == stride, pad 0 (the last layer of the UNET spec, e.g. k32 s32
 * 16x8x8 -> 1x256x256: nn.ConvTranspose2d unet.py:138-140), fused with torch.sigmoid (unet.py:162) and with
 * masked_mse_loss + lambda * pearson (unet.py:314-320,635-678).  Non-overlapping output patches: one tap per input
 * channel and output pixel.
 *   cae_patch_head_fwd : yhat (written only if `yhat` is given: apply / score) and, if target.t0.p is set, the loss:
 *                        loss_out[slot] = masked MSE, pearson_out[slot] = 1 - mean corr, gradient coefficients in
 *                        coef / scalars for the backward call.  Training never writes yhat.
 *   cae_patch_head_bwd : recomputes yhat, forms dL/d(pre-sigmoid) in registers and produces, in one pass over the
 *                        target, partial rows of grad_w [Cin][Cout][K][K] / grad_b [Cout] and the input gradient `din`
 *                        (through the PLAIN / MASK / MASKSTATS epilogue of the producing layer).
 * Workspaces: moments N*Cout*Hin*(K*K/128)*7 doubles (one row per plane, patch row and warp), coef N*Cout*3 floats, scalars 4 floats, partials
 * cae_patch_head_partials_len() floats.  Supported: K in {16, 32}, Cin <= 16, Win <= 64
 * (cae_patch_head_supported); other geometries use cae_conv_up / cae_conv_down / cae_conv_wgrad. */
typedef struct CaePatchHead {
    CaeSrc        in;             /* layer input [N, Cin, Hin, Win] (BatchNorm+ReLU of the producer applied on load) */
    const float*  weight;         /* [Cin][Cout][K][K] */
    const float*  bias;           /* [Cout] or NULL */
    int           K;              /* kernel size == stride */
    int           Cout;
    CaeSrc        target;         /* [N, Cout, K*Hin, K*Win]; t0.p NULL = no loss */
    CaeSrc        mask;           /* [N, 1 or Cout, K*Hin, K*Win]; t0.p NULL = all ones */
    int           mask_channels;
    float         lambda_pearson;
    float         count_scale;    /* as CaeEpilogue.count_scale */
    double*       moments;
    float*        coef;
    float*        scalars;
    float*        loss_out;
    float*        pearson_out;
    unsigned int* ticket;         /* 1 x u32, zero before first use (loss only) */
    const float*  mse_scale;      /* device, may be NULL: as cae_masked_pearson_loss.mse_scale */
} CaePatchHead;
int       cae_patch_head_supported(int K, int stride, int pad, int Cin, int Win);
int       cae_patch_head_fwd(const CaePatchHead* h, const CaeView* yhat, void* stream);
int       cae_patch_head_bwd(const CaePatchHead* h, const CaeView* din, const CaeEpilogue* din_epilogue, float* partials,
                             void* stream);
/* grad_w / grad_b = fixed-order sum of the partial rows cae_patch_head_bwd left in `partials` (only the optimiser waits
 * for it: the engine runs it beside the rest of the backward pass) */
int       cae_patch_head_wgrad_reduce(const CaePatchHead* h, float* grad_w, float* grad_b, const float* partials,
                                      void* stream);
long long cae_patch_head_partials_len(const CaePatchHead* h);

/* ---- variational bottleneck (VarAEModel; the reference names the variant - cli/train_cae.py:32-33,42,
 * model_evaluator.py:35 - but ships no implementation: parity unpinned) --------------------------
 * forward : z = mu + eps*exp(logvar/2) (sample != 0) or z = mu ; kl_out[slot] = kl_scale * KL,
 *           KL = -1/2 * sum_{n,l}(1 + logvar - mu^2 - exp(logvar)) / n_samples
 * backward: dmu = dz + w*mu ; dlogvar = dz*eps*exp(logvar/2)/2 + w*(exp(logvar)-1)/2 , w = kl_weight / n_samples
 * eps is a pre-drawn device array; the batch is selected by cursor[0]*eps_stride like CaeSrc. */
int cae_vae_reparam_fwd(const float* mu, const float* logvar, const float* eps, long long eps_stride,
                        const int* cursor, float* z, int n_samples, int latent, int sample, float kl_scale,
                        float* kl_out, void* stream);
int cae_vae_reparam_bwd(const float* dz, const float* mu, const float* logvar, const float* eps,
                        long long eps_stride, const int* cursor, float* dmu, float* dlogvar, int n_samples,
                        int latent, float kl_weight, void* stream);
/* out[i] ~ N(0,1): counter-based generator keyed by (seed, step_count[0], i); replaces torch.randn_like in the
 * reparameterisation so that a replayed CUDA graph draws fresh noise every step */
int cae_randn(float* out, long long n, unsigned long long seed, const int* step_count, void* stream);
/* out = a + b (gradient fan-in) */
int cae_add2(const float* a, const float* b, float* out, long long n, void* stream);

/* ---- tensor-core GEMM (tc_gemm.cu): tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA ----------
 * C[m,n] = sum_k A[m,k] * B[n,k] for the genuinely dense contractions of the path: the fat ConvTranspose2d layers
 * (decoder.py:44-48 -> aten::convolution / convolution_backward: Cin >= 64) as GEMM + col2im / im2col, and nn.Linear
 * (linear.py:43, unet.py:92-100,121-129).  Operands arrive split for 3xTF32 (x = hi + lo, cae_tc_split or the
 * producers below): hi*hi + hi*lo + lo*hi keeps ~2^-21 relative error per product (the path's 1e-4 bar); a_lo = b_lo =
 * NULL selects 1xTF32.  An operand is K-major (A[m*lda + k]) or MN-major (A[k*lda + m]); pitches are multiples of 4
 * floats, pointers 16-byte aligned.  splits > 1: split-K, slice z writes its partial tile to C + z*split_stride (the
 * caller reduces the slices in a fixed order: cae_tc_reduce).  tile_n: 128 (default) or 256. */
typedef struct CaeTcGemm {
    int M, N, K;
    const float *a_hi, *a_lo;
    long long lda;
    int a_mn_major;
    const float *b_hi, *b_lo;
    long long ldb;
    int b_mn_major;
    float* C;
    long long ldc;
    int splits;
    long long split_stride;
    int tile_n;
} CaeTcGemm;
int cae_tc_gemm(const CaeTcGemm* g, void* stream);
/* hi = x rounded to TF32 (round to nearest), lo = x - hi rounded to TF32 */
int cae_tc_split(const float* x, float* hi, float* lo, long long n, void* stream);

/* ---- ConvTranspose2d on the tensor cores (tc_conv.cu) --------------------------------------------------------------
 * Drop-in replacements of cae_conv_up (forward), cae_conv_down-as-input-gradient and cae_conv_wgrad for a transposed
 * convolution (padding 0) whose channel counts make it a dense contraction (decoder.py:44-48 at config-4 widths):
 * pack -> tcgen05 GEMM (3xTF32) -> col2im / unpack with the conv family's epilogues.  Tensors outside stay fp32 NCHW.
 * Workspaces are caller-owned: a_hi/a_lo [N*Hin*Win, lda] (written by the forward call, read again by the weight
 * gradient), w_hi/w_lo (>= max(kh*kw*Cout*lda, Cin*ldn) floats each), cols (GEMM outputs; cols_len floats:
 * >= N*Hin*Win*max(ldn, lda) and >= splits*Cin*ldn, splits from cae_tc_convt_wgrad_splits), dcols_hi/dcols_lo
 * [N*Hin*Win, ldn] (written by cae_tc_convt_im2col, read by _dgrad and _wgrad). */
typedef struct CaeTcConv {
    int Cin, Cout, kh, kw, stride;
    int N, Hin, Win, Hout, Wout;
    float *a_hi, *a_lo;
    long long lda;          /* >= Cin, multiple of 4 */
    float *w_hi, *w_lo;
    float* cols;
    long long cols_len;
    float *dcols_hi, *dcols_lo;
    long long ldn;          /* >= kh*kw*Cout, multiple of 4 */
} CaeTcConv;
int       cae_tc_convt_supported(int Cin, int Cout, int kh, int kw, int stride, int pad);
long long cae_tc_convt_wgrad_splits(const CaeTcConv* c);
/* y = ConvTranspose2d(transform(in)) + bias; epilogue PLAIN or STATS (BatchNorm statistics, as cae_conv_up) */
int cae_tc_convt_fwd(const CaeTcConv* c, const CaeSrc* in, const float* weight, const CaeView* out, const CaeEpilogue* epi,
                     void* stream);
/* dcols = im2col(transform(dy)) for the two backward GEMMs of the layer */
int cae_tc_convt_im2col(const CaeTcConv* c, const CaeSrc* dy, void* stream);
/* dx = input gradient; epilogue PLAIN, MASK or MASKSTATS (as cae_conv_down used as the input gradient) */
int cae_tc_convt_dgrad(const CaeTcConv* c, const float* weight, const CaeView* dx, const CaeEpilogue* epi, void* stream);
/* grad[Cin][Cout][kh][kw] = weight gradient (split-K, slices summed in index order) */
int cae_tc_convt_wgrad(const CaeTcConv* c, float* grad, void* stream);

/* ---- data ingest on the device (ingest.cu): the host-side steps on either side of the hot path ------
 * cae_minmax: out3 = {min, max, NaN count} of a raw fp32 array (DSDataset.__init__ scan, ds_dataset.py:49-75);
 * partials: cae_minmax_partials_len() floats, ticket: one zeroed uint.
 * cae_normalise_gather: dst[i][:] = (src[order[i]][:] - lo) / (hi - lo) (0 when hi == lo; copy when normalise == 0):
 * min-max normalisation + batch assembly in the shuffled order (ds_dataset.py:99-113,137-159 + default collate), written
 * into a channel slice of the batch tensor (dst_sample_stride >= sample_elems).  order == NULL: identity. */
long long cae_minmax_partials_len(void);
int cae_minmax(const float* x, long long n, float* partials, unsigned int* ticket, float* out3, void* stream);
int cae_normalise_gather(const float* src, long long sample_elems, const int* order, int n_out, float lo, float hi,
                         int normalise, float* dst, long long dst_sample_stride, void* stream);
/* post-training metrics on the device (model_metric.py:19-71 as used by base_model.py:116-125): out[n*8 + {0..7}] = count,
 * sum a, sum e, sum a^2, sum e^2, sum a*e, sum |a-e|, sum (a-e)^2 over the kept pixels of case n (mask > 0; mask == NULL:
 * all), a = actual[n] (raw fp32), e = lo + yhat[n] * scale in float64 (the reference's de-normalisation of its float32
 * predictions, ds_dataset.py:122-125).  mask_per_case elements of mask per case, tiled over the case (1 or C channels). */
int cae_case_metrics(const float* yhat, const float* actual, const float* mask, int n_cases, long long per_case,
                     long long mask_per_case, double lo, double scale, double* out, void* stream);

/* ---- data-parallel exchange fused into the optimiser (dp_fused.cu) -------------------------------------------------------
 * One launch per rank: all-reduce of the flat gradient arena by peer reads over NVLink (summed in rank order: identical
 * bits everywhere) + Adam / AdamW + step bookkeeping.  Replaces NCCL all-reduce + cae_adam_advance for small arenas (the
 * exchange the reference does not have: SURVEY 2.3, 8e).  grads[r] / flags[r]: rank r's gradient arena / flag array
 * (>= 16 uint32, zeroed once) as mapped into THIS process (e.g. torch symmetric memory buffer_ptrs); epoch: a private,
 * zero-initialised device counter of this rank; every rank must call it the same number of times. */
typedef struct CaeDpPeers {
    int world, rank;
    const float* grads[8];
    unsigned int* flags[8];
} CaeDpPeers;
/* first launch of every step: returns (on the stream) once every peer has read this rank's gradients of the previous
 * cae_adam_allreduce - the backward pass may then overwrite them */
int cae_dp_wait_done(const CaeDpPeers* peers, const unsigned int* epoch, void* stream);
int cae_adam_allreduce(float* p, const CaeDpPeers* peers, float* m, float* v, long long n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int decoupled, float grad_scale, int* step_count,
                       int* cursor, int n_batches, unsigned int* epoch, unsigned int* ticket, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAE_B200_H */
