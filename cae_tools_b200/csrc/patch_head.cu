// "Patch head": the last layer of the UNET spec is a transposed convolution whose kernel equals its stride
// (k32 s32 p0: 16x8x8 -> 1x256x256; reference unet.py:138-140 with torch.sigmoid unet.py:162 and the loss
// unet.py:314-320,635-678).  Output patches do not overlap, so every output pixel has exactly one tap per input
// channel:
//   yhat[n,co,K*i+ky,K*j+kx] = sigmoid(b[co] + sum_ci a(n,ci,i,j) * W[ci,co,ky,kx])
// The layer produces 99 % of the bytes of the model, so it is fused end to end:
//   k_ph_fwd    : recomputable forward; optionally writes yhat (apply / score), optionally reads the target (+mask)
//                 and accumulates the seven masked moments per patch row that the loss needs.  In training yhat is
//                 never written: HBM traffic = one read of the target.
//   ph_finalize : moments -> masked MSE, 1 - mean Pearson, per-plane gradient coefficients (last CTA of k_ph_fwd)
//   k_ph_bwd    : recomputes yhat, forms dL/d(pre-sigmoid) in registers and feeds all three consumers at once -
//                 weight gradient (register accumulators, one partial row per CTA), bias gradient and the input
//                 gradient (warp-shuffle transpose reduction) with the ReLU-mask / BatchNorm-backward epilogue of
//                 the previous layer.  HBM traffic = one more read of the target.
//   k_ph_wgrad_reduce: fixed-order sum of the partial rows.
// Thread layout: thread = one group of 4 consecutive taps (ky, kx..kx+3) of the patch; its Cin x 4 weights live in
// registers for the whole kernel; a warp reads/writes 128-byte row segments.  K in {16, 32}, Cin <= 16.
#include "capi_host.h"
#include "conv_family.cuh"

#define PH_CIN 16
// The Pearson statistics are accumulated on (yhat - PH_SHIFT, target - PH_SHIFT): single-pass raw moments of data that live in
// [0, 1] around 0.5 lose ~2 digits in M_dd - mu^2 M (fp32 per-thread partial sums); every gradient behind the head then sat
// 2-4e-5 (max-norm relative) from the float64 evaluation of the step - 20x the fp32 reference's own distance - and the 50-epoch
// loss curve drifted 3x beyond the reference's thread-count spread (tools/grad_deviation.py, profiles/r02_parity_notes.md).
// ph_finalize's algebra is unchanged: fed with shifted moments it returns the gradient coefficients of the shifted variables.
#define PH_SHIFT 0.5f

struct PhArgs {
    CaeSrc in;
    const float* w;
    const float* bias;
    int Cin, Cout, N, Hin, Win;
    CaeView yhat;              // p == NULL: do not write
    CaeSrc target, mask;       // target.t0.p == NULL: no loss; mask.t0.p == NULL: ones
    int mask_channels;
    double* moments;           // [N*Cout][Hin][7]
    float* coef;               // [N*Cout][3]
    float* scalars;            // [0] mse gradient factor, [1] masked mse, [2] pearson term
    float* loss_out;
    float* pearson_out;
    unsigned int* ticket;
    float lambda_pearson, count_scale;
    const float* mse_scale;    // per-batch factor of the masked-MSE term (replaces count_scale there), may be NULL
    int pixels_per_plane;
    // backward
    CaeView dout;              // [N, Cin, Hin, Win]
    CaeEpilogue epi;
    float* partials;           // [rows][Cin*Cout*K*K]
    float* dbpart;             // [rows][Cout]
    float* grad_w;
    float* grad_b;
    int rows;
};

__device__ __forceinline__ float4 ph_ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// sigmoid: cae_fast_sigmoid (common.cuh) - __expf + __frcp_rn still expand to ~16 instructions (the IEEE-rounded
// reciprocal is a Newton iteration, ex2 without .ftz carries denormal scaling): ncu showed 127 instructions per 4-pixel
// strip against 32 FFMA2 of useful work.
__device__ __forceinline__ float ph_sigmoid(float v) { return cae_fast_sigmoid(v); }
// Optional IEEE form for the training kernels (1 / (1 + expf(-v)), what torch evaluates).  Measured (tools/grad_deviation.py):
// no effect on the distance of the gradients from the float64 evaluation, +8 us per step - the deviation came from the
// un-centred loss moments (PH_SHIFT below), not from the SFU approximations.  Kept off.
#ifndef CAE_PH_EXACT_TRAIN_SIGMOID
#define CAE_PH_EXACT_TRAIN_SIGMOID 0
#endif
__device__ __forceinline__ float ph_sigmoid_train(float v) {
#if CAE_PH_EXACT_TRAIN_SIGMOID
    return 1.f / (1.f + expf(-v));
#else
    return cae_fast_sigmoid(v);
#endif
}

// packed fp32 pairs: sm_100 issues two FMAs per FFMA2 instruction (fma.rn.f32x2), which halves the issue slots of the
// three 16 x 4 FMA blocks these kernels are made of
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
struct W4 { f32x2 lo, hi; };          // four taps of one input channel
__device__ __forceinline__ W4 ph_ldw(const float* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    W4 w;
    w.lo = pk(v.x, v.y);
    w.hi = pk(v.z, v.w);
    return w;
}

// stage the activated inputs of patch row (n, i) for one slot, every value DUPLICATED into a pair so that one
// 16-byte shared-memory load yields two ready-made FFMA2 operands: s[(j*PH_CIN + ci)*2 + {0,1}], zero for ci >= Cin
__device__ __forceinline__ void ph_stage(const PhArgs& a, float* s, int n, int i, long long in_base, int tl, int TG) {
    const CaeView& iv = a.in.t0;
    for (int e = tl; e < a.Win * PH_CIN; e += TG) {
        const int j = e / PH_CIN, ci = e - j * PH_CIN;
        float v = 0.f;
        if (ci < a.Cin) {
            const ChanCoef kc = load_coef(a.in, ci);
            v = src_value(a.in, in_base + (long long)n * iv.sN + (long long)ci * iv.sC + (long long)i * iv.ld + j, kc);
        }
        reinterpret_cast<float2*>(s)[e] = make_float2(v, v);
    }
}

// pre-activation of 4 pixels: acc = b + sum_ci a[ci] * w[ci][0..3]; av[ci] = {a, a}
__device__ __forceinline__ void ph_preact(const W4 (&w)[PH_CIN], const float* sa, float b, f32x2 (&av)[PH_CIN], f32x2& acc01,
                                          f32x2& acc23) {
#pragma unroll
    for (int q = 0; q < PH_CIN / 2; ++q) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(sa + 4 * q);
        av[2 * q] = t.x;
        av[2 * q + 1] = t.y;
    }
    // two independent accumulation chains per pair (even / odd channels): the 16-deep FFMA2 chain was the longest
    // fixed-latency dependency of a strip ("wait" is the top stall reason in ncu)
    acc01 = acc23 = pk(b, b);
    f32x2 b01 = 0ull, b23 = 0ull;
#pragma unroll
    for (int ci = 0; ci < PH_CIN; ci += 2) {
        acc01 = fma2(av[ci], w[ci].lo, acc01);
        acc23 = fma2(av[ci], w[ci].hi, acc23);
        b01 = fma2(av[ci + 1], w[ci + 1].lo, b01);
        b23 = fma2(av[ci + 1], w[ci + 1].hi, b23);
    }
    acc01 = add2(acc01, b01);
    acc23 = add2(acc23, b23);
}

// one CTA (the last of k_ph_fwd to finish): per-plane moments (sum of the patch-row partials, in row order) -> losses +
// gradient coefficients.  Same algebra as k_mp_finalize (unet_ops.cuh).
__device__ __forceinline__ void ph_finalize(const PhArgs& a, int rows_per_plane) {
    __shared__ double red[CAE_NWARP];
    __shared__ double s_mo[CAE_NT][7];
    const int C = a.Cout, NC = a.N * C;
    double sq = 0.0, cnt = 0.0, corr_sum = 0.0;
    for (int p0 = 0; p0 < NC; p0 += CAE_NT) {
        const int pn = min(CAE_NT, NC - p0);
        // phase 1 - one thread per (plane, moment): its partial rows are independent loads (all in flight at once: the first
        // version walked the planes warp by warp, eight dependent L2 round trips + shuffle chains = half of the kernel's time,
        // ncu sm__cycles_active max 50.7 k against avg 24.3 k), summed in row order
        __syncthreads();
        for (int e = threadIdx.x; e < pn * 7; e += CAE_NT) {
            const int q = e / 7, k = e - q * 7;
            const double* src = a.moments + (size_t)(p0 + q) * rows_per_plane * 7 + k;
            double t = 0.0;
            int r = 0;
            for (; r + 7 < rows_per_plane; r += 8) {
                double v[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) v[x] = __ldcg(src + (size_t)(r + x) * 7);
#pragma unroll
                for (int x = 0; x < 8; ++x) t += v[x];
            }
            for (; r < rows_per_plane; ++r) t += __ldcg(src + (size_t)r * 7);
            s_mo[q][k] = t;
        }
        __syncthreads();
        // phase 2 - one thread per plane: the (double precision, division / sqrt heavy) algebra runs in parallel
        if ((int)threadIdx.x >= pn) continue;
        const int p = p0 + threadIdx.x;
        const double* mo = s_mo[threadIdx.x];
        const double M = a.mask.t0.p ? mo[0] : (double)a.pixels_per_plane, Md = mo[1], Mt = mo[2], Mdd = mo[3], Mtt = mo[4], Mdt = mo[5];
        sq += mo[6];
        if (a.mask_channels != 1 || (p % C) == 0) cnt += M;
        const double Mp = M + 1e-8;
        const double mu_d = Md / Mp, mu_t = Mt / Mp;
        const double vdd = Mdd - 2.0 * mu_d * Md + mu_d * mu_d * M;
        const double vtt = Mtt - 2.0 * mu_t * Mt + mu_t * mu_t * M;
        const double sd = sqrt(vdd / Mp + 1e-8), st = sqrt(vtt / Mp + 1e-8);
        const double S = (Mdt - mu_t * Md - mu_d * Mt + mu_d * mu_t * M) / st;
        const double T = (Mt - mu_t * M) / st;
        const double D = Md - mu_d * M;
        const double corr = M > 0.0 ? S / (M * sd) : 0.0;
        corr_sum += corr;
        const double wgt = M > 0.0 ? -(double)a.lambda_pearson / ((double)NC * M * sd) : 0.0;
        const double kk = S / (sd * sd * Mp);
        a.coef[p * 3 + 0] = (float)(wgt / st);
        a.coef[p * 3 + 1] = (float)(-wgt * kk);
        a.coef[p * 3 + 2] = (float)(wgt * (-mu_t / st - T / Mp + kk * (mu_d + D / Mp)));
    }
    auto block_sum = [&](double v) {
        v = warp_sum_d(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0.0;
        for (int q = 0; q < CAE_NWARP; ++q) s += red[q];
        return s;
    };
    const double SQ = block_sum(sq), CNT = block_sum(cnt), CS = block_sum(corr_sum);
    if (threadIdx.x == 0) {
        const double cs = a.count_scale > 0.f ? (double)a.count_scale : 1.0;
        const double mse = SQ / CNT;
        const int slot = a.target.cursor ? __ldg(a.target.cursor) : 0;
        // data parallelism with masks: this share's valid pixels / the global batch's (see cae_b200.h)
        const double cm = a.mse_scale ? (double)__ldg(a.mse_scale + slot) : cs;
        a.scalars[0] = (float)(2.0 / CNT * cm);
        a.scalars[1] = (float)mse;
        a.scalars[2] = (float)(1.0 - CS / NC);
        if (a.loss_out) a.loss_out[slot] = (float)(mse * cm);
        if (a.pearson_out) a.pearson_out[slot] = (float)((1.0 - CS / NC) * cs);
    }
}

#define PH_UC 4          // patch rows staged per slot and pass (one exposure of the input-load latency per pass)

// unit (patch row) u of this CTA's contiguous share [ub, ue): slot s takes ub + s, ub + s + SLOTS, ...
struct PhRange { int ub, ue; };
__device__ __forceinline__ PhRange ph_range(int units, int slots) {
    int upc = (units + gridDim.x - 1) / gridDim.x;
    upc = (upc + slots - 1) / slots * slots;
    PhRange r;
    r.ub = blockIdx.x * upc;
    r.ue = min(units, r.ub + upc);
    return r;
}

// Compile-time variants: LOSS (read the target, accumulate the moments) / MASK (a mask tensor is present) / WRITE (store
// yhat); PF = strips in flight per thread (Win % PF == 0).  Runtime flags in the streaming loop cost more issue slots
// than the arithmetic (ncu: 257 instructions per 4-pixel strip, 32 of them the FFMA2 block).
template <int K, bool LOSS, bool MASK, bool WRITE, int PF>
__global__ void __launch_bounds__(CAE_NT, 2) k_ph_fwd(const PhArgs a) {
    constexpr int TG = K * K / 4, SLOTS = CAE_NT / TG, WPS = TG / 32, TPR = K / 4;
    extern __shared__ __align__(16) float s_a[];               // [SLOTS][PH_UC][Win][PH_CIN][2]
    const int tid = threadIdx.x, slot = tid / TG, tl = tid - slot * TG, lane = tid & 31, wis = tl >> 5;
    const int ky = tl / TPR, kx = (tl - ky * TPR) * 4;
    const int co = blockIdx.y;
    W4 w[PH_CIN];
#pragma unroll
    for (int ci = 0; ci < PH_CIN; ++ci) {
        w[ci].lo = w[ci].hi = 0ull;
        if (ci < a.Cin) w[ci] = ph_ldw(a.w + (((size_t)ci * a.Cout + co) * K + ky) * K + kx);
    }
    const float b = a.bias ? __ldg(a.bias + co) : 0.f;
    const long long in_base = src_cursor_offset(a.in);
    const long long tbase = LOSS ? src_cursor_offset(a.target) : 0ll;
    const long long mbase = MASK ? src_cursor_offset(a.mask) : 0ll;
    const CaeView& tv = a.target.t0;
    const CaeView& mv = a.mask.t0;
    const int mc = a.mask_channels == 1 ? 0 : co;
    const PhRange rg = ph_range(a.N * a.Hin, SLOTS);
    const int usz = a.Win * PH_CIN * 2;                         // duplicated pairs
    float* sa0 = s_a + slot * PH_UC * usz;
    // moments of the current plane, per thread, across all of its patch rows that this slot handles (flushed when the
    // plane changes): M (mask only), Md, Mt, Mdd, Mtt, Mdt, E
    float mo[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    __shared__ double s_fl[CAE_NWARP][7];
    auto flush = [&](int n, int i, bool real) {
        if constexpr (SLOTS == 1) {
            // one slot: the CTA's warps all work on this patch row - sum them here (fixed order), one partial row per
            // (plane, patch row): the finalising CTA reads 8x fewer rows
            double* row = a.moments + (((size_t)n * a.Cout + co) * a.Hin + i) * 7;
            if (real) {                                         // (CTA-uniform)
                if (!MASK) mo[0] = 0.f;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const double sk = warp_sum_d((double)mo[k]);
                    if (lane == 0) s_fl[wis][k] = sk;
                    mo[k] = 0.f;
                }
                __syncthreads();
                if (tid < 7) {
                    double t = 0.0;
#pragma unroll
                    for (int q = 0; q < WPS; ++q) t += s_fl[q][tid];
                    row[tid] = t;
                }
                __syncthreads();
            } else if (tid < 7) {
                row[tid] = 0.0;
            }
        } else {
            // one partial row per (plane, patch row, warp); rows of a run other than its last one are written as zeros
            double* row = a.moments + ((((size_t)n * a.Cout + co) * a.Hin + i) * WPS + wis) * 7;
            if (real) {
                if (!MASK) mo[0] = 0.f;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const double sk = warp_sum_d((double)mo[k]);
                    if (lane == 0) row[k] = sk;
                    mo[k] = 0.f;
                }
            } else if (lane < 7) {
                row[lane] = 0.0;
            }
        }
    };
    for (int cb = rg.ub; cb < rg.ue; cb += PH_UC * SLOTS) {
        __syncthreads();
        for (int uu = 0; uu < PH_UC; ++uu) {
            const int u = cb + uu * SLOTS + slot;
            if (u < rg.ue) ph_stage(a, sa0 + uu * usz, u / a.Hin, u % a.Hin, in_base, tl, TG);
        }
        __syncthreads();
        for (int uu = 0; uu < PH_UC; ++uu) {
            const int u = cb + uu * SLOTS + slot;
            if (u >= rg.ue) break;
            const int n = u / a.Hin, i = u - n * a.Hin;
            const float* sa = sa0 + uu * usz;
            const int oy = i * K + ky;
            const float* tp = LOSS ? tv.p + tbase + (long long)n * tv.sN + (long long)co * tv.sC + (long long)oy * tv.ld + kx : nullptr;
            const float* mp = MASK ? mv.p + mbase + (long long)n * mv.sN + (long long)mc * mv.sC + (long long)oy * mv.ld + kx : nullptr;
            float* yp = WRITE ? a.yhat.p + (long long)n * a.yhat.sN + (long long)co * a.yhat.sC + (long long)oy * a.yhat.ld + kx : nullptr;
            for (int j0 = 0; j0 < a.Win; j0 += PF) {
                float4 t4[PF], m4[PF];
#pragma unroll
                for (int q = 0; q < PF; ++q) {
                    if (LOSS) t4[q] = ph_ld4(tp + (j0 + q) * K);
                    if (MASK) m4[q] = ph_ld4(mp + (j0 + q) * K);
                }
#pragma unroll
                for (int q = 0; q < PF; ++q) {
                    const int j = j0 + q;
                    f32x2 av[PH_CIN], acc01, acc23;
                    ph_preact(w, sa + j * PH_CIN * 2, b, av, acc01, acc23);
                    float acc[4], d[4];
                    upk(acc01, acc[0], acc[1]);
                    upk(acc23, acc[2], acc[3]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) d[k] = LOSS ? ph_sigmoid_train(acc[k]) : ph_sigmoid(acc[k]);
                    if (WRITE) __stcs(reinterpret_cast<float4*>(yp + j * K), make_float4(d[0], d[1], d[2], d[3]));
                    if (LOSS) {
                        // first / second moments of (d - PH_SHIFT), (t - PH_SHIFT): variances and the covariance are shift
                        // invariant, and centring the [0, 1] data removes most of the cancellation in M_dd - mu^2 M
                        const float t[4] = {t4[q].x, t4[q].y, t4[q].z, t4[q].w};
                        if (MASK) {
                            const float m[4] = {m4[q].x, m4[q].y, m4[q].z, m4[q].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float ds = d[k] - PH_SHIFT, ts = t[k] - PH_SHIFT;
                                const float md = m[k] * ds, mt = m[k] * ts, e = (d[k] - t[k]) * m[k];
                                mo[0] += m[k]; mo[1] += md; mo[2] += mt;
                                mo[3] = fmaf(md, ds, mo[3]); mo[4] = fmaf(mt, ts, mo[4]); mo[5] = fmaf(md, ts, mo[5]);
                                mo[6] = fmaf(e, e, mo[6]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float ds = d[k] - PH_SHIFT, ts = t[k] - PH_SHIFT;
                                const float e = d[k] - t[k];
                                mo[1] += ds; mo[2] += ts;
                                mo[3] = fmaf(ds, ds, mo[3]); mo[4] = fmaf(ts, ts, mo[4]); mo[5] = fmaf(ds, ts, mo[5]);
                                mo[6] = fmaf(e, e, mo[6]);
                            }
                        }
                    }
                }
            }
            if (LOSS) {
                // last patch row of this plane that this slot handles?
                const int un = u + SLOTS;
                const bool last = un >= rg.ue || un / a.Hin != n;
                flush(n, i, last);
            }
        }
    }
    if (LOSS && cae_last_block(a.ticket)) ph_finalize(a, SLOTS == 1 ? a.Hin : a.Hin * WPS);
}

// 16 per-lane values -> warp sums, one channel per lane pair: lane l ends up with the sum of v[(l >> 1) & 15]
__device__ __forceinline__ float ph_warp_transpose_sum(float (&v)[PH_CIN], int lane) {
    float r8[8], r4[4], r2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float send = hi ? v[k] : v[k + 8], keep = hi ? v[k + 8] : v[k];
            r8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float send = hi ? r8[k] : r8[k + 4], keep = hi ? r8[k + 4] : r8[k];
            r4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float send = hi ? r4[k] : r4[k + 2], keep = hi ? r4[k + 2] : r4[k];
            r2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    const bool hi = lane & 2;
    const float send = hi ? r2[0] : r2[1], keep = hi ? r2[1] : r2[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

template <int K, bool MASK, int PF>
__global__ void __launch_bounds__(CAE_NT, 1) k_ph_bwd(const PhArgs a) {
    constexpr int TG = K * K / 4, SLOTS = CAE_NT / TG, WPS = TG / 32, TPR = K / 4, KK = K * K;
    extern __shared__ __align__(16) float smem[];
    const int usz = a.Win * PH_CIN, usz2 = 2 * usz;
    float* s_a = smem;                                         // [SLOTS][PH_UC][Win][PH_CIN][2] (duplicated pairs)
    float* s_red = smem + SLOTS * PH_UC * usz2;                // [SLOTS][PH_UC][WPS][Win][PH_CIN]
    __shared__ double s_db[CAE_NWARP];
    __shared__ float s_st[CAE_NT][2];
    const int tid = threadIdx.x, slot = tid / TG, tl = tid - slot * TG, lane = tid & 31, wis = tl >> 5;
    const int ky = tl / TPR, kx = (tl - ky * TPR) * 4;
    const long long in_base = src_cursor_offset(a.in);
    const long long tbase = src_cursor_offset(a.target);
    const long long mbase = MASK ? src_cursor_offset(a.mask) : 0ll;
    const CaeView& tv = a.target.t0;
    const CaeView& mv = a.mask.t0;
    const float c0 = a.scalars[0];
    const float cs = a.count_scale > 0.f ? a.count_scale : 1.f;
    float* sa0 = s_a + slot * PH_UC * usz2;
    float* sr0 = s_red + (size_t)slot * PH_UC * WPS * usz;
    const int my_ci = tl & (PH_CIN - 1);                       // channel this thread finishes in the input-gradient tail
    const EpiCh ech = epi_load_channel(a.epi, min(my_ci, a.Cin - 1), my_ci < a.Cin);
    float s1 = 0.f, s2 = 0.f;
    const size_t nelem = (size_t)a.Cin * a.Cout * KK;
    const int row = blockIdx.x * SLOTS + slot;
    const PhRange rg = ph_range(a.N * a.Hin, SLOTS);
    for (int co = 0; co < a.Cout; ++co) {
        W4 w[PH_CIN], gw[PH_CIN];
#pragma unroll
        for (int ci = 0; ci < PH_CIN; ++ci) {
            w[ci].lo = w[ci].hi = 0ull;
            if (ci < a.Cin) w[ci] = ph_ldw(a.w + (((size_t)ci * a.Cout + co) * K + ky) * K + kx);
            gw[ci].lo = gw[ci].hi = 0ull;                      // bit pattern of {0.f, 0.f}
        }
        const float b = a.bias ? __ldg(a.bias + co) : 0.f;
        const int mc = a.mask_channels == 1 ? 0 : co;
        const bool last_co = co == a.Cout - 1;
        double dbs = 0.0;          // the bias gradient is a heavily cancelling sum: fp64 per thread
        for (int cb = rg.ub; cb < rg.ue; cb += PH_UC * SLOTS) {
            __syncthreads();
            for (int uu = 0; uu < PH_UC; ++uu) {
                const int u = cb + uu * SLOTS + slot;
                if (u < rg.ue) ph_stage(a, sa0 + uu * usz2, u / a.Hin, u % a.Hin, in_base, tl, TG);
            }
            __syncthreads();
            for (int uu = 0; uu < PH_UC; ++uu) {
                const int u = cb + uu * SLOTS + slot;
                if (u >= rg.ue) break;
                const int n = u / a.Hin, i = u - n * a.Hin;
                const float* sa = sa0 + uu * usz2;
                float* sr = sr0 + (size_t)(uu * WPS + wis) * usz;
                const int plane = n * a.Cout + co;
                const float ca = a.coef[plane * 3 + 0] * cs, cb2 = a.coef[plane * 3 + 1] * cs, ce = a.coef[plane * 3 + 2] * cs;
                const int oy = i * K + ky;
                const float* tp = tv.p + tbase + (long long)n * tv.sN + (long long)co * tv.sC + (long long)oy * tv.ld + kx;
                const float* mp = MASK ? mv.p + mbase + (long long)n * mv.sN + (long long)mc * mv.sC + (long long)oy * mv.ld + kx
                                       : nullptr;
                // without a mask: g = c0 (d - t) + ca t' + cb d' + ce = gd d' + gt t' + ce   (d' = d - PH_SHIFT, t' = t - PH_SHIFT;
                // the coefficients are those of the shifted variables, see ph_finalize)
                const float gd = c0 + cb2, gt = ca - c0;
                for (int j0 = 0; j0 < a.Win; j0 += PF) {
                    float4 t4[PF], m4[PF];
#pragma unroll
                    for (int q = 0; q < PF; ++q) {
                        t4[q] = ph_ld4(tp + (j0 + q) * K);
                        if (MASK) m4[q] = ph_ld4(mp + (j0 + q) * K);
                    }
#pragma unroll
                    for (int q = 0; q < PF; ++q) {
                        const int j = j0 + q;
                        {
                            f32x2 av[PH_CIN], acc01, acc23;
                            ph_preact(w, sa + j * PH_CIN * 2, b, av, acc01, acc23);
                            float acc[4];
                            upk(acc01, acc[0], acc[1]);
                            upk(acc23, acc[2], acc[3]);
                            const float t[4] = {t4[q].x, t4[q].y, t4[q].z, t4[q].w};
                            float dz[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float d = ph_sigmoid_train(acc[k]);
                                float g;
                                if (MASK) {
                                    const float m = k == 0 ? m4[q].x : (k == 1 ? m4[q].y : (k == 2 ? m4[q].z : m4[q].w));
                                    g = m * fmaf(c0 * m, d - t[k], fmaf(ca, t[k] - PH_SHIFT, fmaf(cb2, d - PH_SHIFT, ce)));
                                } else {
                                    g = fmaf(gd, d - PH_SHIFT, fmaf(gt, t[k] - PH_SHIFT, ce));
                                }
                                dz[k] = g * fmaf(-d, d, d);
                            }
                            dbs += (double)((dz[0] + dz[1]) + (dz[2] + dz[3]));
                            const f32x2 dz01 = pk(dz[0], dz[1]), dz23 = pk(dz[2], dz[3]);
                            float part[PH_CIN];
#pragma unroll
                            for (int ci = 0; ci < PH_CIN; ++ci) {
                                gw[ci].lo = fma2(av[ci], dz01, gw[ci].lo);
                                gw[ci].hi = fma2(av[ci], dz23, gw[ci].hi);
                                float p0, p1;
                                upk(fma2(dz01, w[ci].lo, mul2(dz23, w[ci].hi)), p0, p1);
                                part[ci] = p0 + p1;
                            }
                            const float r = ph_warp_transpose_sum(part, lane);
                            if (!(lane & 1)) sr[j * PH_CIN + ((lane >> 1) & 15)] = r;
                        }
                    }
                }
            }
            __syncthreads();
            // input gradient of this pass: sum the warps' pieces, then the epilogue of the producing layer
            for (int e = tl; e < PH_UC * usz; e += TG) {
                const int uu = e / usz, r2 = e - uu * usz, j = r2 / PH_CIN, ci = r2 - j * PH_CIN;      // ci == my_ci
                const int u = cb + uu * SLOTS + slot;
                if (u < rg.ue && ci < a.Cin) {
                    const int n = u / a.Hin, i = u - n * a.Hin;
                    float v = 0.f;
#pragma unroll
                    for (int q = 0; q < WPS; ++q) v += sr0[(size_t)(uu * WPS + q) * usz + r2];
                    const long long off = (long long)n * a.dout.sN + (long long)ci * a.dout.sC + (long long)i * a.dout.ld + j;
                    if (co > 0) v += a.dout.p[off];
                    if (!last_co) a.dout.p[off] = v;
                    else epi_element(a.epi, a.dout, ech, n, ci, i, j, v, 0ll, 0.f, s1, s2);
                }
            }
        }
        // this slot's weight-gradient row
        if (row < a.rows) {
#pragma unroll
            for (int ci = 0; ci < PH_CIN; ++ci)
                if (ci < a.Cin)
                    *reinterpret_cast<ulonglong2*>(a.partials + (size_t)row * nelem + ((size_t)ci * a.Cout + co) * KK + tl * 4) =
                        make_ulonglong2(gw[ci].lo, gw[ci].hi);
        }
        const double dbw = warp_sum_d(dbs);
        __syncthreads();
        if (lane == 0) s_db[tid >> 5] = dbw;
        __syncthreads();
        if (tl == 0 && row < a.rows) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < WPS; ++q) s += s_db[slot * WPS + q];
            a.dbpart[(size_t)row * a.Cout + co] = (float)s;
        }
    }
    if (epi_reduces(a.epi.mode)) {
        s_st[tid][0] = s1; s_st[tid][1] = s2;
        __syncthreads();
        const int C = a.dout.C;
        if (tid < 2 * PH_CIN) {
            const int c = tid >> 1, st = tid & 1;
            if (c < C) {
                double s = 0.0;
                for (int q = c; q < CAE_NT; q += PH_CIN) s += (double)s_st[q][st];
                a.epi.partials[((size_t)blockIdx.x * C + c) * 2 + st] = s;
            }
        }
        if (cae_last_block(a.epi.ticket)) {
            const double count = (double)a.dout.N * a.dout.H * a.dout.W;
            if (a.epi.mode == CAE_EPI_STATS) finalize_bn_forward(a.epi.bn, a.epi.partials, gridDim.x, count);
            else if (a.epi.mode == CAE_EPI_MASKSTATS) finalize_bn_backward(a.epi.bn, a.epi.partials, gridDim.x, count);
        }
    }
}

// grad_w[e] = sum_rows partials[row][e] (fixed order); CTA = 32 float4 columns x 8 row groups, 8 row loads in flight per
// thread (the rows were just written and sit in L2: the kernel is pure latency - with 64 columns x 4 groups, 4 loads in
// flight and the bias rows summed by ONE thread per channel it took 12 us, 9.6 of them CTA 0's serial bias loop).
// One extra CTA sums the bias rows: a warp per channel, lanes over the rows, fixed-order butterfly in float64.
__global__ void __launch_bounds__(CAE_NT) k_ph_wgrad_reduce(const PhArgs a, long long nelem4) {
    __shared__ float4 s_p[8][32];
    if (blockIdx.x == gridDim.x - 1) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int co = warp; a.grad_b && co < a.Cout; co += CAE_NWARP) {
            double t = 0.0;
            for (int r = lane; r < a.rows; r += 32) t += (double)__ldcg(a.dbpart + (size_t)r * a.Cout + co);
            t = warp_sum_d(t);
            if (lane == 0) a.grad_b[co] = (float)t;
        }
        return;
    }
    const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long long e4 = (long long)blockIdx.x * 32 + col;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e4 < nelem4) {
        const float4* base = reinterpret_cast<const float4*>(a.partials) + e4;
        int r = grp;
        for (; r + 56 < a.rows; r += 64) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldcg(base + (size_t)(r + 8 * j) * nelem4);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
        }
        for (; r < a.rows; r += 8) {
            const float4 v = __ldcg(base + (size_t)r * nelem4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    s_p[grp][col] = s;
    __syncthreads();
    if (grp == 0 && e4 < nelem4) {
        float4 t = s_p[0][col];
#pragma unroll
        for (int g = 1; g < 8; ++g) { t.x += s_p[g][col].x; t.y += s_p[g][col].y; t.z += s_p[g][col].z; t.w += s_p[g][col].w; }
        reinterpret_cast<float4*>(a.grad_w)[e4] = t;
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
static bool ph_plain_src(const CaeSrc& s) { return !s.t1 && !s.k0 && !s.k1 && !s.k2 && !s.relu && !s.kn; }

static int ph_fill(PhArgs& a, const CaePatchHead* h) {
    CAE_REQUIRE(h && h->weight, "patch_head: null argument");
    int rc;
    if ((rc = check_view(h->in.t0, "patch_head input"))) return rc;
    memset(&a, 0, sizeof(a));
    a.in = h->in; a.w = h->weight; a.bias = h->bias;
    a.Cin = h->in.t0.C; a.Cout = h->Cout; a.N = h->in.t0.N; a.Hin = h->in.t0.H; a.Win = h->in.t0.W;
    CAE_REQUIRE(h->K == 16 || h->K == 32, "patch_head: kernel %d not supported (16 or 32)", h->K);
    CAE_REQUIRE(a.Cin <= PH_CIN && a.Cout >= 1 && a.Win <= 64, "patch_head: Cin %d (<= %d) / Cout %d / Win %d (<= 64) not supported",
                a.Cin, PH_CIN, a.Cout, a.Win);
    CAE_REQUIRE(h->in.kn == nullptr, "patch_head: per-(n,c) multipliers are not supported on the input");
    CAE_REQUIRE((uintptr_t)h->weight % 16 == 0, "patch_head: weight must be 16-byte aligned");
    a.target = h->target; a.mask = h->mask; a.mask_channels = h->mask_channels;
    a.moments = h->moments; a.coef = h->coef; a.scalars = h->scalars;
    a.loss_out = h->loss_out; a.pearson_out = h->pearson_out; a.ticket = h->ticket;
    a.lambda_pearson = h->lambda_pearson; a.count_scale = h->count_scale;
    a.mse_scale = h->mse_scale;
    const int Ho = h->K * a.Hin, Wo = h->K * a.Win;
    a.pixels_per_plane = Ho * Wo;
    if (a.target.t0.p) {
        const CaeView& t = a.target.t0;
        CAE_REQUIRE(t.N == a.N && t.C == a.Cout && t.H == Ho && t.W == Wo, "patch_head: target geometry %dx%dx%dx%d != %dx%dx%dx%d",
                    t.N, t.C, t.H, t.W, a.N, a.Cout, Ho, Wo);
        CAE_REQUIRE(ph_plain_src(a.target) && src_aligned(a.target), "patch_head: target must be a plain, 16-byte aligned tensor");
        CAE_REQUIRE(a.moments && a.coef && a.scalars && a.ticket, "patch_head: loss workspace missing");
        if (a.mask.t0.p) {
            const CaeView& m = a.mask.t0;
            CAE_REQUIRE((a.mask_channels == 1 || a.mask_channels == a.Cout) && m.C == a.mask_channels && m.N == a.N &&
                            m.H == Ho && m.W == Wo, "patch_head: mask must be [N, 1 or C, H, W]");
            CAE_REQUIRE(ph_plain_src(a.mask) && src_aligned(a.mask), "patch_head: mask must be a plain, 16-byte aligned tensor");
        } else {
            a.mask_channels = a.Cout;
        }
    }
    return CAE_OK;
}

// CTAs: every CTA gets the same number of patch rows (a multiple of the slot count), at most one CTA per SM
static int ph_grid(const PhArgs& a, int K, int per_sm = 1) {
    const int slots = CAE_NT / (K * K / 4);
    const int units = a.N * a.Hin;
    int gx = min(ceil_div(units, slots), CAE_NUM_SMS * per_sm);
    int upc = ceil_div(units, gx);
    upc = ceil_div(upc, slots) * slots;
    return ceil_div(units, upc);
}

extern "C" int cae_patch_head_supported(int K, int stride, int pad, int Cin, int Win) {
    return (K == stride && pad == 0 && (K == 16 || K == 32) && Cin <= PH_CIN && Win <= 64) ? 1 : 0;
}

extern "C" int cae_patch_head_fwd(const CaePatchHead* h, const CaeView* yhat, void* stream) {
    PhArgs a;
    int rc = ph_fill(a, h);
    if (rc) return rc;
    const int K = h->K;
    if (yhat && yhat->p) {
        CAE_REQUIRE(yhat->N == a.N && yhat->C == a.Cout && yhat->H == K * a.Hin && yhat->W == K * a.Win && view_aligned(*yhat),
                    "patch_head_fwd: yhat geometry / alignment");
        a.yhat = *yhat;
    }
    CAE_REQUIRE(a.yhat.p || a.target.t0.p, "patch_head_fwd: nothing to do (no yhat, no target)");
    cudaStream_t st = (cudaStream_t)stream;
    const int slots = CAE_NT / (K * K / 4);
    const size_t smem = (size_t)slots * PH_UC * a.Win * PH_CIN * 2 * 4;
    dim3 grid(ph_grid(a, K, 2), a.Cout);
    const bool loss = a.target.t0.p != nullptr, mask = loss && a.mask.t0.p != nullptr, write = a.yhat.p != nullptr;
    CAE_REQUIRE(!(loss && write), "patch_head_fwd: writing yhat and computing the loss in one call is not supported "
                                  "(score writes, train / test reduce)");
    const bool pf2 = a.Win % 2 == 0, pf4 = write && a.Win % 4 == 0;     // the write-only variant has registers to spare
#define PH_FWD(K_, L_, M_, W_)                                                                      \
    do {                                                                                            \
        if (W_ && pf4) k_ph_fwd<K_, L_, M_, W_, (W_ ? 4 : 2)><<<grid, CAE_NT, smem, st>>>(a);       \
        else if (pf2) k_ph_fwd<K_, L_, M_, W_, 2><<<grid, CAE_NT, smem, st>>>(a);                   \
        else k_ph_fwd<K_, L_, M_, W_, 1><<<grid, CAE_NT, smem, st>>>(a);                            \
    } while (0)
    if (K == 32) {
        if (write) PH_FWD(32, false, false, true);
        else if (mask) PH_FWD(32, true, true, false);
        else PH_FWD(32, true, false, false);
    } else {
        if (write) PH_FWD(16, false, false, true);
        else if (mask) PH_FWD(16, true, true, false);
        else PH_FWD(16, true, false, false);
    }
#undef PH_FWD
    return cae_check_launch("cae_patch_head_fwd");
}

extern "C" long long cae_patch_head_partials_len(const CaePatchHead* h) {
    PhArgs a;
    if (ph_fill(a, h)) return -1;
    const int slots = CAE_NT / (h->K * h->K / 4);
    const long long rows = (long long)CAE_NUM_SMS * slots;
    return rows * ((long long)a.Cin * a.Cout * h->K * h->K + a.Cout);
}

extern "C" int cae_patch_head_bwd(const CaePatchHead* h, const CaeView* din, const CaeEpilogue* epi, float* partials,
                                  void* stream) {
    PhArgs a;
    int rc = ph_fill(a, h);
    if (rc) return rc;
    CAE_REQUIRE(a.target.t0.p, "patch_head_bwd: needs the target");
    CAE_REQUIRE(din && epi && partials, "patch_head_bwd: null argument");
    CAE_REQUIRE(din->p && din->N == a.N && din->C == a.Cin && din->H == a.Hin && din->W == a.Win,
                "patch_head_bwd: input-gradient geometry differs from the input");
    CAE_REQUIRE((uintptr_t)partials % 16 == 0, "patch_head_bwd: partials must be 16-byte aligned");
    a.dout = *din;
    a.epi = *epi;
    if (a.epi.mode == CAE_EPI_MASKSTATS && a.epi.act.p == nullptr) a.epi.mode = CAE_EPI_PLAIN;
    CAE_REQUIRE(a.epi.mode == CAE_EPI_PLAIN || a.epi.mode == CAE_EPI_MASK || a.epi.mode == CAE_EPI_MASKSTATS,
                "patch_head_bwd: epilogue mode %d not supported", a.epi.mode);
    CAE_REQUIRE(a.epi.addend.t0.p == nullptr && a.epi.bias == nullptr, "patch_head_bwd: addend / bias not supported");
    if (a.epi.mode != CAE_EPI_PLAIN) {
        const CaeView& v = a.epi.act;
        CAE_REQUIRE(v.p && v.N == a.N && v.C == a.Cin && v.H == a.Hin && v.W == a.Win, "patch_head_bwd: act geometry");
    }
    if (a.epi.mode == CAE_EPI_MASKSTATS) {
        CAE_REQUIRE(a.epi.partials && a.epi.ticket && a.epi.bn.C == a.Cin && a.epi.bn.scale && a.epi.bn.shift && a.epi.bn.mean &&
                        a.epi.bn.invstd && a.epi.bn.bwdA && a.epi.bn.bwdB && a.epi.bn.bwdC, "patch_head_bwd: BN block incomplete");
    }
    const int K = h->K;
    const int slots = CAE_NT / (K * K / 4), wps = K * K / 128;
    const int gx = ph_grid(a, K);
    a.rows = gx * slots;
    const long long nelem = (long long)a.Cin * a.Cout * K * K;
    a.partials = partials;
    a.dbpart = partials + (long long)CAE_NUM_SMS * slots * nelem;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)slots * PH_UC * a.Win * PH_CIN * 4 * (2 + wps);
    const bool mask = a.mask.t0.p != nullptr, pf4 = a.Win % 2 == 0;     // two strips in flight (four spill at 255 registers)
#define PH_BWD(K_, M_, P_)                                        \
    do {                                                          \
        ensure_smem(k_ph_bwd<K_, M_, P_>);                        \
        k_ph_bwd<K_, M_, P_><<<gx, CAE_NT, smem, st>>>(a);        \
    } while (0)
#define PH_BWD_K(K_)                                              \
    do {                                                          \
        if (mask) { if (pf4) PH_BWD(K_, true, 2); else PH_BWD(K_, true, 1); }     \
        else { if (pf4) PH_BWD(K_, false, 2); else PH_BWD(K_, false, 1); }        \
    } while (0)
    if (K == 32) PH_BWD_K(32); else PH_BWD_K(16);
#undef PH_BWD_K
#undef PH_BWD
    return cae_check_launch("cae_patch_head_bwd");
}

extern "C" int cae_patch_head_wgrad_reduce(const CaePatchHead* h, float* grad_w, float* grad_b, const float* partials,
                                           void* stream) {
    PhArgs a;
    int rc = ph_fill(a, h);
    if (rc) return rc;
    CAE_REQUIRE(grad_w && partials && (uintptr_t)grad_w % 16 == 0 && (uintptr_t)partials % 16 == 0,
                "patch_head_wgrad_reduce: null / misaligned argument");
    const int K = h->K;
    const int slots = CAE_NT / (K * K / 4);
    a.rows = ph_grid(a, K) * slots;
    const long long nelem = (long long)a.Cin * a.Cout * K * K;
    a.partials = const_cast<float*>(partials);
    a.dbpart = a.partials + (long long)CAE_NUM_SMS * slots * nelem;
    a.grad_w = grad_w; a.grad_b = grad_b;
    k_ph_wgrad_reduce<<<ceil_div(nelem / 4, 32) + 1, CAE_NT, 0, (cudaStream_t)stream>>>(a, nelem / 4);
    return cae_check_launch("cae_patch_head_wgrad_reduce");
}
