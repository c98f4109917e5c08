// Fused UNET decoder block around the ChannelAttention gate (reference unet.py:23-39 ChannelAttention, :149-163
// Decoder.forward: x = convT(x); x = x * attention(x); x = cat(x, skip); x = relu(bn(x))).
// The tensors between the transposed convolutions are tiny (<= 16x8x8 per sample in the shipped spec), so the chain
//   plane statistics -> 1x1 MLP -> gate -> concat with the skip -> BatchNorm statistics            (forward)
//   dL/d att -> MLP backward -> gate backward -> bias-gradient sums                                 (backward)
// is latency-, not bandwidth-bound: five launches per block cost 5 x (launch + L2 round trips).  Here each direction is
// ONE kernel, one CTA per sample (grid-stride over samples), the sample's planes staged in shared memory;
// cross-sample sums (BatchNorm statistics, dW1, dW2, dbias) go through per-CTA partial rows and a last-CTA fixed-order
// reduction like every other reduction in the library.
#include "capi_host.h"

#define AB_MAX_ELEMS 10240          // C*H*W floats of one sample kept in shared memory (40 KB)
#define AB_MAX_C 128
#define AB_MAX_ROWS 296

struct AbFwdArgs {
    CaeView y;                 // transposed-conv output (bias included) [N, C, H, W]
    CaeSrc skip;               // encoder activation read through BN+ReLU on load, [N, C, H, W]
    const float* W1;           // [Cr][C]
    const float* W2;           // [C][Cr]
    int Cr;
    CaeView cat;               // [N, 2C, H, W]
    int train;                 // accumulate BatchNorm statistics of cat
    CaeBN bn;                  // BatchNorm2d(2C)
    double* partials;          // [rows][2C][2]
    unsigned int* ticket;
    float* stats;              // [N*C][4]: sum, sum sq, max, argmax
    float* att;                // [N][C]
    float* hid;                // [N][2][Cr]
};

__global__ void __launch_bounds__(CAE_NT) k_att_block_fwd(const AbFwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int C = a.y.C, H = a.y.H, W = a.y.W, HW = H * W, Cr = a.Cr;
    float* ys = sm;                                   // [C][HW]
    float* avg = ys + C * HW;                         // [C]
    float* mx = avg + C;                              // [C]
    float* h = mx + C;                                // [2][Cr]
    float* at = h + 2 * Cr;                           // [C]
    double* acc = reinterpret_cast<double*>(sm + ((C * HW + 3 * C + 2 * Cr + 1) & ~1));   // [2C][2], across this CTA's samples
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long skip_base = src_cursor_offset(a.skip);
    const CaeView& sv = a.skip.t0;
    for (int i = tid; i < 4 * C; i += CAE_NT) acc[i] = 0.0;
    for (int n = blockIdx.x; n < a.y.N; n += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < C * HW; e += CAE_NT) {
            const int c = e / HW, r = e - c * HW, yy = r / W, xx = r - yy * W;
            ys[e] = __ldg(a.y.p + (long long)n * a.y.sN + (long long)c * a.y.sC + (long long)yy * a.y.ld + xx);
        }
        __syncthreads();
        // plane statistics: one warp per channel; arg-max = first occurrence (torch.max semantics of AdaptiveMaxPool2d)
        for (int c = warp; c < C; c += CAE_NWARP) {
            float s = 0.f, q = 0.f, m = -INFINITY;
            int am = 0x7fffffff;
            for (int i = lane; i < HW; i += 32) {
                const float v = ys[c * HW + i];
                s += v;
                q = fmaf(v, v, q);
                if (v > m) { m = v; am = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
                const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
                if (m2 > m || (m2 == m && a2 < am)) { m = m2; am = a2; }
            }
            const double S = warp_sum_d((double)s), Q = warp_sum_d((double)q);
            if (lane == 0) {
                avg[c] = (float)(S / HW);
                mx[c] = m;
                float* st = a.stats + ((size_t)n * C + c) * 4;
                st[0] = (float)S; st[1] = (float)Q; st[2] = m; st[3] = (float)am;
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * Cr; i += CAE_NT) {
            const int which = i / Cr, r = i - which * Cr;
            const float* src = which ? mx : avg;
            float v = 0.f;
            for (int c = 0; c < C; ++c) v = fmaf(__ldg(a.W1 + r * C + c), src[c], v);
            v = fmaxf(v, 0.f);
            h[i] = v;
            a.hid[(size_t)n * 2 * Cr + i] = v;
        }
        __syncthreads();
        for (int c = tid; c < C; c += CAE_NT) {
            float v = 0.f;
            for (int r = 0; r < Cr; ++r) v = fmaf(__ldg(a.W2 + c * Cr + r), h[r] + h[Cr + r], v);
            v = 1.f / (1.f + expf(-v));
            at[c] = v;
            a.att[(size_t)n * C + c] = v;
        }
        __syncthreads();
        // cat = [att * y ; skip], BatchNorm statistics per channel (one warp per channel)
        for (int c2 = warp; c2 < 2 * C; c2 += CAE_NWARP) {
            float s = 0.f, q = 0.f;
            float* ob = a.cat.p + (long long)n * a.cat.sN + (long long)c2 * a.cat.sC;
            if (c2 < C) {
                const float g = at[c2];
                for (int i = lane; i < HW; i += 32) {
                    const int yy = i / W, xx = i - yy * W;
                    const float v = g * ys[c2 * HW + i];
                    ob[(long long)yy * a.cat.ld + xx] = v;
                    s += v;
                    q = fmaf(v, v, q);
                }
            } else {
                const int c = c2 - C;
                const ChanCoef kc = load_coef(a.skip, c);
                for (int i = lane; i < HW; i += 32) {
                    const int yy = i / W, xx = i - yy * W;
                    const float v = src_value(a.skip, skip_base + (long long)n * sv.sN + (long long)c * sv.sC +
                                                          (long long)yy * sv.ld + xx, kc);
                    ob[(long long)yy * a.cat.ld + xx] = v;
                    s += v;
                    q = fmaf(v, v, q);
                }
            }
            if (a.train) {
                const double S = warp_sum_d((double)s), Q = warp_sum_d((double)q);
                if (lane == 0) { acc[2 * c2] += S; acc[2 * c2 + 1] += Q; }
            }
        }
    }
    if (a.train) {
        __syncthreads();
        for (int i = tid; i < 4 * C; i += CAE_NT) a.partials[(size_t)blockIdx.x * 4 * C + i] = acc[i];
        if (cae_last_block(a.ticket))
            finalize_bn_forward(a.bn, a.partials, gridDim.x, (double)a.y.N * HW);
    }
}

struct AbBwdArgs {
    CaeSrc g;                  // dL/d(att*y): gated half of the concat gradient through the BN-backward affine
    CaeView y;
    const float* att;
    const float* hid;
    const float* stats;
    const float* W1;
    const float* W2;
    int Cr;
    CaeView dy;                // dL/dy [N, C, H, W]
    float* dW1;                // [Cr][C]
    float* dW2;                // [C][Cr]
    float* dbias;              // [C] : sum over (n, h, w) of dy
    float* partials;           // [rows][2*Cr*C + C]
    unsigned int* ticket;
};

__global__ void __launch_bounds__(CAE_NT) k_att_block_bwd(const AbBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int C = a.y.C, H = a.y.H, W = a.y.W, HW = H * W, Cr = a.Cr;
    const int NW = 2 * Cr * C + C;
    float* gs = sm;                       // [C][HW]
    float* ds = gs + C * HW;              // [C]
    float* avg = ds + C;                  // [C]
    float* mx = avg + C;                  // [C]
    float* dav = mx + C;                  // [C] dL/d avg per pixel
    float* dmx = dav + C;                 // [C]
    float* dh = dmx + C;                  // [2][Cr]
    float* hh = dh + 2 * Cr;              // [2][Cr]
    float* accw = hh + 2 * Cr;            // [NW]: dW1 | dW2 | dbias over this CTA's samples
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CaeView& gv = a.g.t0;
    const long long gbase = src_cursor_offset(a.g);
    const float inv_hw = 1.f / (float)HW;
    for (int i = tid; i < NW; i += CAE_NT) accw[i] = 0.f;
    for (int n = blockIdx.x; n < a.y.N; n += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < C * HW; e += CAE_NT) {
            const int c = e / HW, r = e - c * HW, yy = r / W, xx = r - yy * W;
            const ChanCoef kc = load_coef(a.g, c);
            gs[e] = src_value(a.g, gbase + (long long)n * gv.sN + (long long)c * gv.sC + (long long)yy * gv.ld + xx, kc);
        }
        for (int i = tid; i < 2 * Cr; i += CAE_NT) hh[i] = a.hid[(size_t)n * 2 * Cr + i];
        __syncthreads();
        // dL/d att[c] = sum_hw g * y ; ds = datt * att * (1 - att)
        for (int c = warp; c < C; c += CAE_NWARP) {
            float s = 0.f;
            const float* yb = a.y.p + (long long)n * a.y.sN + (long long)c * a.y.sC;
            for (int i = lane; i < HW; i += 32) {
                const int yy = i / W, xx = i - yy * W;
                s = fmaf(gs[c * HW + i], __ldg(yb + (long long)yy * a.y.ld + xx), s);
            }
            const double S = warp_sum_d((double)s);
            if (lane == 0) {
                const float at = a.att[(size_t)n * C + c];
                ds[c] = (float)S * at * (1.f - at);
                avg[c] = a.stats[((size_t)n * C + c) * 4 + 0] * inv_hw;
                mx[c] = a.stats[((size_t)n * C + c) * 4 + 2];
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * Cr; i += CAE_NT) {
            const int r = i % Cr;
            float v = 0.f;
            for (int c = 0; c < C; ++c) v = fmaf(__ldg(a.W2 + c * Cr + r), ds[c], v);
            dh[i] = hh[i] > 0.f ? v : 0.f;
        }
        // dW2[c][r] += ds[c] * (h_avg[r] + h_max[r])
        for (int i = tid; i < C * Cr; i += CAE_NT) {
            const int c = i / Cr, r = i - c * Cr;
            accw[Cr * C + i] = fmaf(ds[c], hh[r] + hh[Cr + r], accw[Cr * C + i]);
        }
        __syncthreads();
        for (int i = tid; i < Cr * C; i += CAE_NT) {
            const int r = i / C, c = i - r * C;
            accw[i] = fmaf(dh[r], avg[c], fmaf(dh[Cr + r], mx[c], accw[i]));
        }
        for (int c = tid; c < C; c += CAE_NT) {
            float ga = 0.f, gm = 0.f;
            for (int r = 0; r < Cr; ++r) {
                const float w = __ldg(a.W1 + r * C + c);
                ga = fmaf(w, dh[r], ga);
                gm = fmaf(w, dh[Cr + r], gm);
            }
            dav[c] = ga * inv_hw;
            dmx[c] = gm;
        }
        __syncthreads();
        // dy = att * g + davg + dmax * [pixel == argmax]; plane sums -> bias gradient
        for (int c = warp; c < C; c += CAE_NWARP) {
            const float at = a.att[(size_t)n * C + c], da = dav[c], dm = dmx[c];
            const int amax = (int)a.stats[((size_t)n * C + c) * 4 + 3];
            float* ob = a.dy.p + (long long)n * a.dy.sN + (long long)c * a.dy.sC;
            float s = 0.f;
            for (int i = lane; i < HW; i += 32) {
                const int yy = i / W, xx = i - yy * W;
                float v = fmaf(at, gs[c * HW + i], da);
                if (i == amax) v += dm;
                ob[(long long)yy * a.dy.ld + xx] = v;
                s += v;
            }
            const double S = warp_sum_d((double)s);
            if (lane == 0) accw[2 * Cr * C + c] += (float)S;
        }
    }
    __syncthreads();
    for (int i = tid; i < NW; i += CAE_NT) a.partials[(size_t)blockIdx.x * NW + i] = accw[i];
    if (cae_last_block(a.ticket)) {
        for (int i = tid; i < NW; i += CAE_NT) {
            float s = 0.f;
            for (int r = 0; r < (int)gridDim.x; ++r) s += __ldcg(a.partials + (size_t)r * NW + i);
            if (i < Cr * C) a.dW1[i] = s;
            else if (i < 2 * Cr * C) a.dW2[i - Cr * C] = s;
            else if (a.dbias) a.dbias[i - 2 * Cr * C] = s;
        }
    }
}

// =====================================================================================================
extern "C" int cae_attention_block_supported(int C, int H, int W, int Cr) {
    if (C < 1 || C > AB_MAX_C || Cr < 1 || H < 1 || W < 1) return 0;
    return ((long long)C * H * W + 2ll * Cr * C + 6 * C + 4 * Cr <= AB_MAX_ELEMS) ? 1 : 0;     // shared memory of the backward kernel
}

extern "C" long long cae_attention_block_partials_len(int C, int Cr) { return (long long)AB_MAX_ROWS * (2 * Cr * C + C); }

static int ab_rows(int N) { return N < AB_MAX_ROWS ? N : AB_MAX_ROWS; }

extern "C" int cae_attention_block_fwd(const CaeView* y, const CaeSrc* skip, const float* W1, const float* W2, int Cr,
                                       const CaeView* cat, const CaeEpilogue* epi, float* stats, float* att, float* hid,
                                       void* stream) {
    CAE_REQUIRE(y && skip && W1 && W2 && cat && epi && stats && att && hid, "attention_block_fwd: null argument");
    int rc;
    if ((rc = check_view(*y, "attention_block_fwd y"))) return rc;
    if ((rc = check_view(skip->t0, "attention_block_fwd skip"))) return rc;
    if ((rc = check_view(*cat, "attention_block_fwd cat"))) return rc;
    CAE_REQUIRE(cae_attention_block_supported(y->C, y->H, y->W, Cr), "attention_block_fwd: %dx%dx%d (Cr %d) too large for the fused kernel",
                y->C, y->H, y->W, Cr);
    const CaeView& s = skip->t0;
    CAE_REQUIRE(s.N == y->N && s.C == y->C && s.H == y->H && s.W == y->W, "attention_block_fwd: skip geometry %dx%dx%dx%d != y %dx%dx%dx%d",
                s.N, s.C, s.H, s.W, y->N, y->C, y->H, y->W);
    CAE_REQUIRE(cat->N == y->N && cat->C == 2 * y->C && cat->H == y->H && cat->W == y->W, "attention_block_fwd: cat must be [N, 2C, H, W]");
    CAE_REQUIRE(skip->kn == nullptr, "attention_block_fwd: kn not supported on the skip operand");
    CAE_REQUIRE(epi->mode == CAE_EPI_PLAIN || epi->mode == CAE_EPI_STATS, "attention_block_fwd: epilogue must be PLAIN or STATS");
    AbFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.y = *y; a.skip = *skip; a.W1 = W1; a.W2 = W2; a.Cr = Cr; a.cat = *cat;
    a.stats = stats; a.att = att; a.hid = hid;
    if (epi->mode == CAE_EPI_STATS) {
        CAE_REQUIRE(epi->partials && epi->ticket && epi->bn.C == 2 * y->C && epi->bn.scale && epi->bn.shift && epi->bn.mean &&
                        epi->bn.invstd, "attention_block_fwd: BN block incomplete (needs C = %d)", 2 * y->C);
        a.train = 1; a.bn = epi->bn; a.partials = epi->partials; a.ticket = epi->ticket;
    }
    const int C = y->C, HW = y->H * y->W;
    const size_t smem = (size_t)((C * HW + 3 * C + 2 * Cr + 1) & ~1) * 4 + (size_t)4 * C * 8;
    // training: one partial row per CTA (<= AB_MAX_ROWS, CTAs loop over samples); eval: no cross-sample state at all, so
    // one CTA per sample up to 16 waves of the machine (apply() batches of 4096 ran 14 samples per CTA in sequence)
    const int rows = a.train ? min(ab_rows(y->N), CAE_MAX_GRID_X) : min(y->N, CAE_NUM_SMS * 8 * 16);
    k_att_block_fwd<<<rows, CAE_NT, smem, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_attention_block_fwd");
}

extern "C" int cae_attention_block_bwd(const CaeSrc* g, const CaeView* y, const float* att, const float* hid,
                                       const float* stats, const float* W1, const float* W2, int Cr, const CaeView* dy,
                                       float* dW1, float* dW2, float* dbias, float* partials, unsigned int* ticket,
                                       void* stream) {
    CAE_REQUIRE(g && y && att && hid && stats && W1 && W2 && dy && dW1 && dW2 && partials && ticket,
                "attention_block_bwd: null argument");
    int rc;
    if ((rc = check_view(*y, "attention_block_bwd y"))) return rc;
    if ((rc = check_view(g->t0, "attention_block_bwd g"))) return rc;
    if ((rc = check_view(*dy, "attention_block_bwd dy"))) return rc;
    CAE_REQUIRE(cae_attention_block_supported(y->C, y->H, y->W, Cr), "attention_block_bwd: %dx%dx%d (Cr %d) too large for the fused kernel",
                y->C, y->H, y->W, Cr);
    const CaeView& gv = g->t0;
    CAE_REQUIRE(gv.N == y->N && gv.C == y->C && gv.H == y->H && gv.W == y->W && dy->N == y->N && dy->C == y->C &&
                    dy->H == y->H && dy->W == y->W, "attention_block_bwd: geometry mismatch");
    CAE_REQUIRE(g->kn == nullptr, "attention_block_bwd: kn not supported");
    AbBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.g = *g; a.y = *y; a.att = att; a.hid = hid; a.stats = stats; a.W1 = W1; a.W2 = W2; a.Cr = Cr; a.dy = *dy;
    a.dW1 = dW1; a.dW2 = dW2; a.dbias = dbias; a.partials = partials; a.ticket = ticket;
    const int C = y->C, HW = y->H * y->W;
    const size_t smem = (size_t)(C * HW + 5 * C + 4 * Cr + 2 * Cr * C + C) * 4;
    k_att_block_bwd<<<ab_rows(y->N), CAE_NT, smem, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_attention_block_bwd");
}
