// Kernels specific to the UNET variant (reference: src/cae_tools/models/unet.py):
//   ChannelAttention gate (unet.py:23-39), skip concatenation (unet.py:149-163), masked MSE + Pearson loss
//   (unet.py:635-678).  Plane = one (n, c) image of H x W pixels.
#pragma once
#include "common.cuh"

// block-wide sum of a double, result in thread 0 (fixed order)
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    __syncthreads();
    return t;
}

// ---------------------------------------------------------------------------------------
// per-plane statistics of y: out[(n*C+c)*4 + {0,1,2,3}] = sum, sum of squares, max, argmax (pixel index as float)
// grid = N*C planes, one CTA each.  (AdaptiveAvgPool2d(1) / AdaptiveMaxPool2d(1) of unet.py:26-27)
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_plane_stats(const CaeView y, float* __restrict__ out) {
    __shared__ double red[CAE_NWARP];
    __shared__ float smax[CAE_NWARP];
    __shared__ int sarg[CAE_NWARP];
    const int plane = blockIdx.x, n = plane / y.C, c = plane - n * y.C;
    const float* base = y.p + (long long)n * y.sN + (long long)c * y.sC;
    const int HW = y.H * y.W;
    double s = 0.0, q = 0.0;
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int i = threadIdx.x; i < HW; i += CAE_NT) {
        const int r = i / y.W, x = i - r * y.W;
        const float v = __ldg(base + (long long)r * y.ld + x);
        s += v;
        q += (double)v * v;
        if (v > mx) { mx = v; am = i; }      // first occurrence within this thread's (increasing) indices
    }
    // warp arg-max with smallest index on ties (torch returns the first maximum)
    for (int o = 16; o > 0; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        int a2 = __shfl_xor_sync(0xffffffffu, am, o);
        if (m2 > mx || (m2 == mx && a2 < am)) { mx = m2; am = a2; }
    }
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; sarg[threadIdx.x >> 5] = am; }
    double S = block_sum_d(s, red);
    double Q = block_sum_d(q, red);
    if (threadIdx.x == 0) {
        float m = smax[0];
        int a = sarg[0];
        for (int w = 1; w < CAE_NWARP; ++w)
            if (smax[w] > m || (smax[w] == m && sarg[w] < a)) { m = smax[w]; a = sarg[w]; }
        out[plane * 4 + 0] = (float)S;
        out[plane * 4 + 1] = (float)Q;
        out[plane * 4 + 2] = m;
        out[plane * 4 + 3] = (float)a;
    }
}

// ---------------------------------------------------------------------------------------
// ChannelAttention forward, one CTA per sample:
//   att = sigmoid( W2 relu(W1 avg) + W2 relu(W1 max) ),  W1 [Cr][C], W2 [C][Cr]   (1x1 convs without bias)
// hid[n][0][r] = relu(W1 avg)_r, hid[n][1][r] = relu(W1 max)_r are kept for the backward pass.
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_ca_fwd(const float* __restrict__ stats, const float* __restrict__ W1,
                                                    const float* __restrict__ W2, int C, int Cr, float inv_hw,
                                                    float* __restrict__ att, float* __restrict__ hid) {
    extern __shared__ float sm[];                 // avg[C], mx[C], h[2*Cr]
    float* avg = sm;
    float* mx = sm + C;
    float* h = sm + 2 * C;
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        avg[c] = stats[(n * C + c) * 4 + 0] * inv_hw;
        mx[c] = stats[(n * C + c) * 4 + 2];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Cr; i += blockDim.x) {
        const int which = i / Cr, r = i - which * Cr;
        const float* src = which ? mx : avg;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(__ldg(W1 + r * C + c), src[c], acc);
        acc = fmaxf(acc, 0.f);
        h[i] = acc;
        hid[(size_t)n * 2 * Cr + i] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < Cr; ++r) acc = fmaf(__ldg(W2 + c * Cr + r), h[r] + h[Cr + r], acc);
        att[n * C + c] = 1.f / (1.f + expf(-acc));
    }
}

// ---------------------------------------------------------------------------------------
// ChannelAttention backward, ONE CTA (deterministic accumulation over the samples):
//   in : datt[N][C] (dL/d att), att, hid, stats (avg/max), W1, W2
//   out: dW1[Cr][C], dW2[C][Cr], davg[N][C] (already divided by H*W: gradient per pixel), dmax[N][C]
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_ca_bwd(const float* __restrict__ datt, const float* __restrict__ att,
                                                    const float* __restrict__ hid, const float* __restrict__ stats,
                                                    const float* __restrict__ W1, const float* __restrict__ W2, int N, int C,
                                                    int Cr, float inv_hw, float* __restrict__ dW1, float* __restrict__ dW2,
                                                    float* __restrict__ davg, float* __restrict__ dmax) {
    extern __shared__ float sm[];                 // ds[C], dh[2*Cr], avg[C], mx[C]
    float* ds = sm;
    float* dh = sm + C;
    float* avg = dh + 2 * Cr;
    float* mx = avg + C;
    for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) { dW1[i] = 0.f; dW2[i] = 0.f; }
    __syncthreads();
    for (int n = 0; n < N; ++n) {
        const float* h = hid + (size_t)n * 2 * Cr;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float a = att[n * C + c];
            ds[c] = datt[n * C + c] * a * (1.f - a);
            avg[c] = stats[(n * C + c) * 4 + 0] * inv_hw;
            mx[c] = stats[(n * C + c) * 4 + 2];
        }
        __syncthreads();
        // dW2[c][r] += ds[c] * (h_avg[r] + h_max[r])   (each element owned by one thread: fixed order over n)
        for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
            const int c = i / Cr, r = i - c * Cr;
            dW2[i] = fmaf(ds[c], h[r] + h[Cr + r], dW2[i]);
        }
        // dh[which][r] = (W2^T ds)_r masked by relu
        for (int i = threadIdx.x; i < 2 * Cr; i += blockDim.x) {
            const int r = i % Cr;
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc = fmaf(__ldg(W2 + c * Cr + r), ds[c], acc);
            dh[i] = h[i] > 0.f ? acc : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) {
            const int r = i / C, c = i - r * C;
            dW1[i] = fmaf(dh[r], avg[c], fmaf(dh[Cr + r], mx[c], dW1[i]));
        }
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float ga = 0.f, gm = 0.f;
            for (int r = 0; r < Cr; ++r) {
                const float w = __ldg(W1 + r * C + c);
                ga = fmaf(w, dh[r], ga);
                gm = fmaf(w, dh[Cr + r], gm);
            }
            davg[n * C + c] = ga * inv_hw;
            dmax[n * C + c] = gm;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// per-plane dot product: out[n*C+c] = sum_hw g(n,c,hw) * y(n,c,hw), g read through an on-load transform
// (dL/d att of the gate g = att * y)
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_plane_dot(const CaeSrc g, const CaeView y, float* __restrict__ out) {
    __shared__ double red[CAE_NWARP];
    const CaeView& gv = g.t0;
    const int plane = blockIdx.x, n = plane / y.C, c = plane - n * y.C;
    const ChanCoef kc = load_coef(g, c);
    const long long gb = (long long)n * gv.sN + (long long)c * gv.sC;
    const float* yb = y.p + (long long)n * y.sN + (long long)c * y.sC;
    const int HW = y.H * y.W;
    double s = 0.0;
    for (int i = threadIdx.x; i < HW; i += CAE_NT) {
        const int r = i / y.W, x = i - r * y.W;
        s += (double)(src_value(g, gb + (long long)r * gv.ld + x, kc) * __ldg(yb + (long long)r * y.ld + x));
    }
    double S = block_sum_d(s, red);
    if (threadIdx.x == 0) out[plane] = (float)S;
}

// ---------------------------------------------------------------------------------------
// dL/dy of a gated transposed-conv output: dy = att * g + davg + dmax * [pixel == argmax]
// (g = dL/d(att*y) through an on-load transform); also the plane sums of dy (bias gradient pieces).
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_gate_bwd(const CaeSrc g, const float* __restrict__ att,
                                                      const float* __restrict__ davg, const float* __restrict__ dmax,
                                                      const float* __restrict__ stats, const CaeView dy,
                                                      float* __restrict__ plane_sum) {
    __shared__ double red[CAE_NWARP];
    const CaeView& gv = g.t0;
    const int plane = blockIdx.x, n = plane / dy.C, c = plane - n * dy.C;
    const ChanCoef kc = load_coef(g, c);
    const float a = att[plane], da = davg[plane], dm = dmax[plane];
    const int amax = (int)stats[plane * 4 + 3];
    const long long gb = (long long)n * gv.sN + (long long)c * gv.sC;
    float* ob = dy.p + (long long)n * dy.sN + (long long)c * dy.sC;
    const int HW = dy.H * dy.W;
    double s = 0.0;
    for (int i = threadIdx.x; i < HW; i += CAE_NT) {
        const int r = i / dy.W, x = i - r * dy.W;
        float v = fmaf(a, src_value(g, gb + (long long)r * gv.ld + x, kc), da);
        if (i == amax) v += dm;
        ob[(long long)r * dy.ld + x] = v;
        s += v;
    }
    double S = block_sum_d(s, red);
    if (threadIdx.x == 0 && plane_sum) plane_sum[plane] = (float)S;
}

// out[c] = sum_n in[n*C + c]  (fixed order)
static __global__ void k_sum_over_n(const float* __restrict__ in, int N, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += in[n * C + c];
        out[c] = s;
    }
}

// ---------------------------------------------------------------------------------------
// masked MSE + Pearson loss (unet.py:635-678), three launches:
//   1. k_mp_moments : per plane  M = sum m, sum m d, sum m t, sum m d^2, sum m t^2, sum m d t, sum (m (d-t))^2
//   2. k_mp_finalize: loss values and the per-plane coefficients of dL/dd = c0 m^2 (d-t) + m (ca t + cb d + ce)
//   3. k_mp_grad    : dz = dL/dd * d (1-d)   (d = sigmoid output), plane sums of dz
// mask has Cm = 1 or C channels (or is absent: all ones).
// ---------------------------------------------------------------------------------------
struct MaskedPearsonArgs {
    CaeView pred;            // d = yhat (sigmoid output)
    CaeSrc target;           // t (batch cursor honoured)
    CaeSrc mask;             // m; mask.t0.p == NULL -> all ones
    int mask_channels;       // 1 or C
    double* moments;         // [N*C][7]
    float* coef;             // [N*C][3]  (ca, cb, ce)
    float* scalars;          // [0] = c0 (mse gradient factor), [1] = masked mse, [2] = pearson loss
    float* loss_out;         // loss_out[slot] = mse (the value the reference's history records)
    float* pearson_out;      // pearson_out[slot] = 1 - mean corr
    float lambda_pearson;
    float count_scale;
    const float* mse_scale;  // per-batch factor of the masked-MSE term (replaces count_scale there), may be NULL
};

__device__ __forceinline__ float mp_mask(const MaskedPearsonArgs& a, long long mbase, int n, int c, int r, int x) {
    if (!a.mask.t0.p) return 1.f;
    const CaeView& mv = a.mask.t0;
    const int mc = a.mask_channels == 1 ? 0 : c;
    return __ldg(mv.p + mbase + (long long)n * mv.sN + (long long)mc * mv.sC + (long long)r * mv.ld + x);
}

static __global__ void __launch_bounds__(CAE_NT) k_mp_moments(const MaskedPearsonArgs a) {
    __shared__ double red[CAE_NWARP];
    const CaeView& pv = a.pred;
    const CaeView& tv = a.target.t0;
    const int plane = blockIdx.x, n = plane / pv.C, c = plane - n * pv.C;
    const long long tbase = src_cursor_offset(a.target), mbase = a.mask.t0.p ? src_cursor_offset(a.mask) : 0ll;
    const int HW = pv.H * pv.W;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < HW; i += CAE_NT) {
        const int r = i / pv.W, x = i - r * pv.W;
        const double d = __ldg(pv.p + (long long)n * pv.sN + (long long)c * pv.sC + (long long)r * pv.ld + x);
        const double t = __ldg(tv.p + tbase + (long long)n * tv.sN + (long long)c * tv.sC + (long long)r * tv.ld + x);
        const double m = mp_mask(a, mbase, n, c, r, x);
        acc[0] += m; acc[1] += m * d; acc[2] += m * t; acc[3] += m * d * d; acc[4] += m * t * t; acc[5] += m * d * t;
        const double e = (d - t) * m;
        acc[6] += e * e;
    }
    for (int k = 0; k < 7; ++k) {
        double S = block_sum_d(acc[k], red);
        if (threadIdx.x == 0) a.moments[(size_t)plane * 7 + k] = S;
    }
}

// single CTA
static __global__ void __launch_bounds__(CAE_NT) k_mp_finalize(const MaskedPearsonArgs a) {
    __shared__ double red[CAE_NWARP];
    const int NC = a.pred.N * a.pred.C, C = a.pred.C;
    double sq = 0.0, cnt = 0.0, corr_sum = 0.0;
    for (int p = threadIdx.x; p < NC; p += CAE_NT) {
        const double* mo = a.moments + (size_t)p * 7;
        const double M = mo[0], Md = mo[1], Mt = mo[2], Mdd = mo[3], Mtt = mo[4], Mdt = mo[5];
        sq += mo[6];
        // count = torch.sum(mask) over the mask tensor as given (not broadcast over channels)
        if (a.mask_channels != 1 || (p % C) == 0) cnt += M;
        const double Mp = M + 1e-8;
        const double mu_d = Md / Mp, mu_t = Mt / Mp;
        // sum m (d-mu_d)^2 etc. from the raw moments
        const double vdd = Mdd - 2.0 * mu_d * Md + mu_d * mu_d * M;
        const double vtt = Mtt - 2.0 * mu_t * Mt + mu_t * mu_t * M;
        const double sd = sqrt(vdd / Mp + 1e-8), st = sqrt(vtt / Mp + 1e-8);
        const double S = (Mdt - mu_t * Md - mu_d * Mt + mu_d * mu_t * M) / st;     // sum m d_c t_hat
        const double T = (Mt - mu_t * M) / st;                                      // sum m t_hat
        const double D = Md - mu_d * M;                                             // sum m d_c
        const double corr = M > 0.0 ? S / (M * sd) : 0.0;
        corr_sum += corr;
        // d corr / d d_i = m_i / (M sd) * [ t_hat_i - T/Mp - S/(sd^2 Mp) (d_c,i - D/Mp) ]
        const double w = M > 0.0 ? -(double)a.lambda_pearson / ((double)NC * M * sd) : 0.0;   // d(1 - mean corr)
        const double kk = S / (sd * sd * Mp);
        a.coef[p * 3 + 0] = (float)(w / st);
        a.coef[p * 3 + 1] = (float)(-w * kk);
        a.coef[p * 3 + 2] = (float)(w * (-mu_t / st - T / Mp + kk * (mu_d + D / Mp)));
    }
    double SQ = block_sum_d(sq, red);
    double CNT = block_sum_d(cnt, red);
    double CS = block_sum_d(corr_sum, red);
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

static __global__ void __launch_bounds__(CAE_NT) k_mp_grad(const MaskedPearsonArgs a, const CaeView dz, float* __restrict__ plane_sum) {
    __shared__ double red[CAE_NWARP];
    const CaeView& pv = a.pred;
    const CaeView& tv = a.target.t0;
    const int plane = blockIdx.x, n = plane / pv.C, c = plane - n * pv.C;
    const long long tbase = src_cursor_offset(a.target), mbase = a.mask.t0.p ? src_cursor_offset(a.mask) : 0ll;
    const float c0 = a.scalars[0];
    const float cs = a.count_scale > 0.f ? a.count_scale : 1.f;
    const float ca = a.coef[plane * 3 + 0] * cs, cb = a.coef[plane * 3 + 1] * cs, ce = a.coef[plane * 3 + 2] * cs;
    const int HW = pv.H * pv.W;
    double s = 0.0;
    for (int i = threadIdx.x; i < HW; i += CAE_NT) {
        const int r = i / pv.W, x = i - r * pv.W;
        const float d = __ldg(pv.p + (long long)n * pv.sN + (long long)c * pv.sC + (long long)r * pv.ld + x);
        const float t = __ldg(tv.p + tbase + (long long)n * tv.sN + (long long)c * tv.sC + (long long)r * tv.ld + x);
        const float m = mp_mask(a, mbase, n, c, r, x);
        const float g = c0 * m * m * (d - t) + m * (ca * t + cb * d + ce);
        const float v = g * d * (1.f - d);
        dz.p[(long long)n * dz.sN + (long long)c * dz.sC + (long long)r * dz.ld + x] = v;
        s += v;
    }
    double S = block_sum_d(s, red);
    if (threadIdx.x == 0 && plane_sum) plane_sum[plane] = (float)S;
}
