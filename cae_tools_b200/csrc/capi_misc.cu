// C ABI, dense / optimiser / loss / UNET / variational entry points.  See include/cae_b200.h.
#include "capi_host.h"
#include "dense_misc.cuh"
#include "unet_ops.cuh"

extern "C" long long cae_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(CaeView);
        case 1: return sizeof(CaeSrc);
        case 2: return sizeof(CaeConvGeom);
        case 3: return sizeof(CaeBN);
        case 4: return sizeof(CaeEpilogue);
        case 5: return sizeof(CaeGemm);
        case 6: return sizeof(CaePatchHead);
        case 7: return sizeof(CaeFcStack);
        case 8: return sizeof(CaeUnetStem);
        case 9: return sizeof(CaeTcGemm);
        case 10: return sizeof(CaeTcConv);
        case 11: return sizeof(CaeStemTrain);
        case 12: return sizeof(CaeDpPeers);
        default: return -1;
    }
}

// ---- dense / misc ------------------------------------------------------------------------
extern "C" int cae_gemm(const CaeGemm* g, void* stream) {
    CAE_REQUIRE(g && g->A && g->B && g->C, "gemm: null argument");
    CAE_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm: empty problem %dx%dx%d", g->M, g->N, g->K);
    CAE_REQUIRE((!g->a_k0 || (g->a_k2 && g->a_hw > 0)) && (!g->b_k0 || (g->b_k2 && g->b_hw > 0)),
                "gemm: on-load affine needs k0, k2 and hw");
    const bool plain = !g->a_k0 && !g->b_k0 && !g->a_relu && !g->b_relu && !g->rowsum_A;
    if (plain && g->K >= 256 && (long long)g->M * g->N <= 16384) {
        k_gemm_warpk<<<ceil_div((long long)g->M * g->N, CAE_NWARP), CAE_NT, 0, (cudaStream_t)stream>>>(*g);
        return cae_check_launch("cae_gemm(warp-k)");
    }
    if (plain && (long long)g->M * g->N <= 8192) {
        k_gemm_skinny<<<ceil_div((long long)g->M * g->N, CAE_NT), CAE_NT, 0, (cudaStream_t)stream>>>(*g);
        return cae_check_launch("cae_gemm(skinny)");
    }
    dim3 grid(ceil_div(g->N, GT), ceil_div(g->M, GT));
    k_gemm<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(*g);
    return cae_check_launch("cae_gemm");
}

extern "C" int cae_bn_eval_prepare(const CaeBN* device_table, int count, void* stream) {
    CAE_REQUIRE(device_table && count > 0, "bn_eval_prepare: bad argument");
    k_bn_eval_prepare<<<count, 128, 0, (cudaStream_t)stream>>>(device_table, count);
    return cae_check_launch("cae_bn_eval_prepare");
}

extern "C" int cae_mse(const float* a, const float* b, long long n, double* partials, unsigned int* ticket,
                       float* loss_out, const int* cursor, void* stream) {
    CAE_REQUIRE(a && b && partials && ticket && loss_out && n > 0, "mse: bad argument");
    int grid = min(ceil_div(n, CAE_NT * 4), CAE_MAX_GRID_X);
    k_mse<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(a, b, n, partials, ticket, loss_out, cursor);
    return cae_check_launch("cae_mse");
}

extern "C" int cae_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int decoupled, float grad_scale, const int* step_count,
                        void* stream) {
    CAE_REQUIRE(p && g && m && v && step_count && n > 0, "adam: bad argument");
    int grid = min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 8);
    k_adam<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled,
                                                       grad_scale, step_count);
    return cae_check_launch("cae_adam");
}

extern "C" int cae_adam_advance(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                float eps, float weight_decay, int decoupled, float grad_scale, int* step_count, int* cursor,
                                int n_batches, unsigned int* ticket, void* stream) {
    CAE_REQUIRE(p && g && m && v && step_count && ticket && n > 0, "adam_advance: bad argument");
    int grid = min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 8);
    k_adam_advance<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled,
                                                               grad_scale, step_count, cursor, n_batches, ticket);
    return cae_check_launch("cae_adam_advance");
}

extern "C" int cae_step_advance(int* step_count, int* cursor, int n_batches, void* stream) {
    CAE_REQUIRE(step_count || cursor, "step_advance: nothing to do");
    k_step_advance<<<1, 32, 0, (cudaStream_t)stream>>>(step_count, cursor, n_batches);
    return cae_check_launch("cae_step_advance");
}

// ---- variational bottleneck ----------------------------------------------------------------------
extern "C" int cae_vae_reparam_fwd(const float* mu, const float* logvar, const float* eps, long long eps_stride,
                                   const int* cursor, float* z, int n_samples, int latent, int sample, float kl_scale,
                                   float* kl_out, void* stream) {
    CAE_REQUIRE(mu && logvar && z && n_samples > 0 && latent > 0, "vae_reparam_fwd: bad argument");
    CAE_REQUIRE(!sample || eps, "vae_reparam_fwd: sampling needs eps");
    k_vae_reparam_fwd<<<1, CAE_NT, 0, (cudaStream_t)stream>>>(mu, logvar, eps, eps_stride, cursor, z,
                                                               n_samples * latent, sample, kl_scale / (float)n_samples,
                                                               kl_out);
    return cae_check_launch("cae_vae_reparam_fwd");
}

extern "C" int cae_vae_reparam_bwd(const float* dz, const float* mu, const float* logvar, const float* eps,
                                   long long eps_stride, const int* cursor, float* dmu, float* dlogvar, int n_samples,
                                   int latent, float kl_weight, void* stream) {
    CAE_REQUIRE(dz && mu && logvar && eps && dmu && dlogvar && n_samples > 0 && latent > 0, "vae_reparam_bwd: bad argument");
    int n = n_samples * latent;
    k_vae_reparam_bwd<<<min(ceil_div(n, CAE_NT), CAE_NUM_SMS), CAE_NT, 0, (cudaStream_t)stream>>>(
        dz, mu, logvar, eps, eps_stride, cursor, dmu, dlogvar, n, kl_weight / (float)n_samples);
    return cae_check_launch("cae_vae_reparam_bwd");
}

extern "C" int cae_add2(const float* a, const float* b, float* out, long long n, void* stream) {
    CAE_REQUIRE(a && b && out && n > 0, "add2: bad argument");
    k_add2<<<min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 4), CAE_NT, 0, (cudaStream_t)stream>>>(a, b, out, n);
    return cae_check_launch("cae_add2");
}

extern "C" int cae_randn(float* out, long long n, unsigned long long seed, const int* step_count, void* stream) {
    CAE_REQUIRE(out && n > 0, "randn: bad argument");
    k_randn<<<min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 4), CAE_NT, 0, (cudaStream_t)stream>>>(out, n, seed, step_count);
    return cae_check_launch("cae_randn");
}

// ---- UNET pieces ---------------------------------------------------------------------------------------
extern "C" int cae_plane_stats(const CaeView* y, float* stats, void* stream) {
    CAE_REQUIRE(y && stats, "plane_stats: null argument");
    int rc = check_view(*y, "plane_stats input");
    if (rc) return rc;
    k_plane_stats<<<y->N * y->C, CAE_NT, 0, (cudaStream_t)stream>>>(*y, stats);
    return cae_check_launch("cae_plane_stats");
}

extern "C" int cae_channel_attention_fwd(const float* stats, const float* W1, const float* W2, int N, int C, int Cr, int HW,
                                         float* att, float* hid, void* stream) {
    CAE_REQUIRE(stats && W1 && W2 && att && hid && N > 0 && C > 0 && Cr > 0 && HW > 0, "channel_attention_fwd: bad argument");
    size_t smem = (size_t)(2 * C + 2 * Cr) * 4;
    CAE_REQUIRE(smem <= 48 * 1024, "channel_attention_fwd: %d channels do not fit", C);
    k_ca_fwd<<<N, CAE_NT, smem, (cudaStream_t)stream>>>(stats, W1, W2, C, Cr, 1.f / (float)HW, att, hid);
    return cae_check_launch("cae_channel_attention_fwd");
}

extern "C" int cae_channel_attention_bwd(const float* datt, const float* att, const float* hid, const float* stats,
                                         const float* W1, const float* W2, int N, int C, int Cr, int HW, float* dW1,
                                         float* dW2, float* davg, float* dmax, void* stream) {
    CAE_REQUIRE(datt && att && hid && stats && W1 && W2 && dW1 && dW2 && davg && dmax && N > 0 && C > 0 && Cr > 0,
                "channel_attention_bwd: bad argument");
    size_t smem = (size_t)(3 * C + 2 * Cr) * 4;
    CAE_REQUIRE(smem <= 48 * 1024, "channel_attention_bwd: %d channels do not fit", C);
    k_ca_bwd<<<1, CAE_NT, smem, (cudaStream_t)stream>>>(datt, att, hid, stats, W1, W2, N, C, Cr, 1.f / (float)HW, dW1,
                                                          dW2, davg, dmax);
    return cae_check_launch("cae_channel_attention_bwd");
}

extern "C" int cae_plane_dot(const CaeSrc* g, const CaeView* y, float* out, void* stream) {
    CAE_REQUIRE(g && y && out, "plane_dot: null argument");
    CAE_REQUIRE(g->t0.N == y->N && g->t0.C == y->C && g->t0.H == y->H && g->t0.W == y->W, "plane_dot: geometry mismatch");
    k_plane_dot<<<y->N * y->C, CAE_NT, 0, (cudaStream_t)stream>>>(*g, *y, out);
    return cae_check_launch("cae_plane_dot");
}

extern "C" int cae_gate_bwd(const CaeSrc* g, const float* att, const float* davg, const float* dmax, const float* stats,
                            const CaeView* dy, float* plane_sum, void* stream) {
    CAE_REQUIRE(g && att && davg && dmax && stats && dy, "gate_bwd: null argument");
    CAE_REQUIRE(g->t0.N == dy->N && g->t0.C == dy->C && g->t0.H == dy->H && g->t0.W == dy->W, "gate_bwd: geometry mismatch");
    k_gate_bwd<<<dy->N * dy->C, CAE_NT, 0, (cudaStream_t)stream>>>(*g, att, davg, dmax, stats, *dy, plane_sum);
    return cae_check_launch("cae_gate_bwd");
}

extern "C" int cae_sum_over_n(const float* in, int N, int C, float* out, void* stream) {
    CAE_REQUIRE(in && out && N > 0 && C > 0, "sum_over_n: bad argument");
    k_sum_over_n<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(in, N, C, out);
    return cae_check_launch("cae_sum_over_n");
}

extern "C" int cae_masked_pearson_loss(const CaeView* pred, const CaeSrc* target, const CaeSrc* mask, int mask_channels,
                                       float lambda_pearson, float count_scale, double* moments, float* coef,
                                       float* scalars, float* loss_out, float* pearson_out, const CaeView* dz,
                                       float* plane_sum, const float* mse_scale, void* stream) {
    CAE_REQUIRE(pred && target && moments && coef && scalars, "masked_pearson_loss: null argument");
    int rc = check_view(*pred, "masked_pearson_loss pred");
    if (rc) return rc;
    const CaeView& t = target->t0;
    CAE_REQUIRE(t.p && t.N == pred->N && t.C == pred->C && t.H == pred->H && t.W == pred->W,
                "masked_pearson_loss: target geometry differs from prediction");
    MaskedPearsonArgs a;
    memset(&a, 0, sizeof(a));
    a.pred = *pred;
    a.target = *target;
    if (mask && mask->t0.p) {
        a.mask = *mask;
        CAE_REQUIRE((mask_channels == 1 || mask_channels == pred->C) && mask->t0.C == mask_channels &&
                        mask->t0.H == pred->H && mask->t0.W == pred->W && mask->t0.N == pred->N,
                    "masked_pearson_loss: mask must be [N, 1 or C, H, W]");
        a.mask_channels = mask_channels;
    } else {
        a.mask_channels = pred->C;
    }
    a.moments = moments; a.coef = coef; a.scalars = scalars;
    a.loss_out = loss_out; a.pearson_out = pearson_out;
    a.lambda_pearson = lambda_pearson; a.count_scale = count_scale;
    a.mse_scale = mse_scale;
    cudaStream_t st = (cudaStream_t)stream;
    const int planes = pred->N * pred->C;
    k_mp_moments<<<planes, CAE_NT, 0, st>>>(a);
    k_mp_finalize<<<1, CAE_NT, 0, st>>>(a);
    if (dz) {
        CAE_REQUIRE(dz->p && dz->N == pred->N && dz->C == pred->C && dz->H == pred->H && dz->W == pred->W,
                    "masked_pearson_loss: dz geometry differs from prediction");
        k_mp_grad<<<planes, CAE_NT, 0, st>>>(a, *dz, plane_sum);
    }
    return cae_check_launch("cae_masked_pearson_loss");
}
