// Shared device/host helpers for libcae_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cae_b200.h"

#define CAE_NT 256            // threads per CTA for all family kernels
#define CAE_NWARP (CAE_NT / 32)
#define CAE_NUM_SMS 148       // B200
#define CAE_MAX_GRID_X (CAE_NUM_SMS * 4)   // persistent-style cap: partial-sum rows per kernel

void cae_set_error(const char* fmt, ...);
int  cae_check_launch(const char* what);

#define CAE_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            cae_set_error(__VA_ARGS__);                \
            return CAE_EINVAL;                         \
        }                                              \
    } while (0)

// ---------------------------------------------------------------------------------------
// operand access
// ---------------------------------------------------------------------------------------
struct ChanCoef {
    float k0, k1, k2;
};

__device__ __forceinline__ ChanCoef load_coef(const CaeSrc& s, int c) {
    ChanCoef k;
    k.k0 = s.k0 ? __ldg(s.k0 + c) : 1.f;
    k.k1 = s.k1 ? __ldg(s.k1 + c) : 0.f;
    k.k2 = s.k2 ? __ldg(s.k2 + c) : 0.f;
    return k;
}

__device__ __forceinline__ long long src_cursor_offset(const CaeSrc& s) {
    return s.cursor ? (long long)(__ldg(s.cursor)) * s.cursor_stride : 0ll;
}

// value of operand element at element offset `off` (cursor offset already included)
__device__ __forceinline__ float src_value(const CaeSrc& s, long long off, const ChanCoef& k) {
    float v = fmaf(__ldg(s.t0.p + off), k.k0, k.k2);
    if (s.t1) v = fmaf(__ldg(s.t1 + off), k.k1, v);
    if (s.relu) v = fmaxf(v, 0.f);
    return v;
}

// sigmoid in four SFU / FMA-pipe instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP); relative error ~2e-7 for |v| < 80.
// Used by the HBM-streaming last-layer kernels, where the IEEE expf + division pair (~25 instructions per pixel) or even
// __expf + __frcp_rn (~16) dominate the issue slots.
__device__ __forceinline__ float cae_fast_sigmoid(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}

// Operand split for the tensor-core paths: x = hi + lo with hi = x ROUNDED to TF32 (round-to-nearest, so |x - hi| <= 2^-11 |x|
// instead of 2^-10 with truncation) and lo = (x - hi) itself rounded to TF32 (the tensor core would otherwise truncate its low
// bits).  What the pair drops is then <= 2^-22 |x|, unbiased; the three products hi*hi + hi*lo + lo*hi (tc_gemm.cu) leave out
// lo*lo <= 2^-22 of the product.  Measured: a 3xTF32 GEMM with chunked accumulation sits 4 - 7e-7 of the max-norm from
// float64, flat in K (profiles/r02_tc_precision.md).
__device__ __forceinline__ float tf32_rn(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) {
    hi = tf32_rn(v);
    lo = tf32_rn(v - hi);
}

// packed fp32 pairs: one FFMA2 (fma.rn.f32x2) issues two IEEE FMAs.  The tile kernels of the wide thin layers
// (conv_wgrad_tile / conv_up_tile / conv_down_tile) pair two adjacent channels, which arrive as one 8-byte shared-memory
// load, and so halve the FMA issue slots of their inner loops (bit-identical to scalar fmaf).
__device__ __forceinline__ unsigned long long wgt_pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void wgt_upk(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long wgt_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------
// deterministic two-stage reduction: every CTA writes one row of partial sums, the CTA that
// takes the last ticket reduces the rows in index order (result independent of scheduling).
// ---------------------------------------------------------------------------------------
// returns true (uniformly over the CTA) in exactly one CTA of the grid: the last to arrive
__device__ __forceinline__ bool cae_last_block(unsigned int* ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == total - 1u);
        if (s_last) *ticket = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Reduce NV per-thread floats over the CTA and store them (as doubles) at dst[0..nvalid).
// Must be called by all CTA_NT threads.
template <int NV>
__device__ __forceinline__ void cta_reduce_store(float (&v)[NV], double* dst, int nvalid) {
    __shared__ double red[CAE_NWARP][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double d = warp_sum_d((double)v[i]);
        if (lane == 0) red[warp][i] = d;
    }
    __syncthreads();
    if ((int)threadIdx.x < nvalid) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < nw; ++w) s += red[w][threadIdx.x];
        dst[threadIdx.x] = s;
    }
    __syncthreads();
}

// Sum column `col` (of `ncols`) over `rows` rows of a row-major double matrix, by one warp,
// in a fixed order; result valid in all lanes.
__device__ __forceinline__ double warp_colsum(const double* part, int rows, int ncols, int col) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int r = lane; r < rows; r += 32) s += __ldcg(part + (size_t)r * ncols + col);
    return warp_sum_d(s);
}

// Both statistics of channel c (columns 2c, 2c+1) over `rows` partial rows by one warp.  All loads of a lane are
// issued before the first add (rows <= CAE_MAX_GRID_X -> at most 19 per lane), so the L2 latency is paid once.
__device__ __forceinline__ void warp_colsum2(const double* part, int rows, int C, int c, double& S, double& Q) {
    constexpr int MAXIT = (CAE_MAX_GRID_X + 31) / 32;
    const int lane = threadIdx.x & 31;
    double2 v[MAXIT];
#pragma unroll
    for (int i = 0; i < MAXIT; ++i) {
        const int r = lane + 32 * i;
        v[i] = (r < rows) ? __ldcg(reinterpret_cast<const double2*>(part + ((size_t)r * C + c) * 2)) : make_double2(0.0, 0.0);
    }
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int i = 0; i < MAXIT; ++i) {
        s += v[i].x;
        q += v[i].y;
    }
    S = warp_sum_d(s);
    Q = warp_sum_d(q);
}

// ---- finalizers (executed by the last CTA; `rows` = gridDim.x partial rows of C*2 doubles) ----
// Column sums of the partial rows for every channel at once: `tpc` threads (a power of two <= 32, groups aligned
// inside a warp) share one channel, each adds the rows r = sub, sub+tpc, ... (loads issued four at a time so the
// L2 latency overlaps), then a fixed-order butterfly combines the group.  f(c, S, Q) runs in the group leader.
template <typename F>
__device__ __forceinline__ void for_each_channel_sums(const double* part, int rows, int C, F f, int ncols = -1) {
    // C: channels per partial row (row stride); ncols: how many channels (from `part` on) to process
    const int NC = ncols < 0 ? C : ncols;
    int tpc = 1;
    while (tpc < 32 && tpc * 2 * NC <= (int)blockDim.x) tpc *= 2;
    const int cpp = blockDim.x / tpc;          // channels per pass
    const int sub = threadIdx.x & (tpc - 1);
    for (int c0 = 0; c0 < NC; c0 += cpp) {
        const int c = c0 + (threadIdx.x / tpc);
        double s = 0.0, q = 0.0;
        if (c < NC) {
            const double2* col = reinterpret_cast<const double2*>(part) + c;
            int r = sub;
            for (; r + 3 * tpc < rows; r += 4 * tpc) {
                double2 v0 = __ldcg(col + (size_t)r * C), v1 = __ldcg(col + (size_t)(r + tpc) * C);
                double2 v2 = __ldcg(col + (size_t)(r + 2 * tpc) * C), v3 = __ldcg(col + (size_t)(r + 3 * tpc) * C);
                s += v0.x; q += v0.y; s += v1.x; q += v1.y; s += v2.x; q += v2.y; s += v3.x; q += v3.y;
            }
            for (; r < rows; r += tpc) {
                double2 v = __ldcg(col + (size_t)r * C);
                s += v.x; q += v.y;
            }
        }
        for (int o = tpc >> 1; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (c < NC && sub == 0) f(c, s, q);
    }
}

// BatchNorm forward statistics (training): reference semantics of nn.BatchNorm2d
// (biased variance for normalisation, unbiased for running_var, momentum update).
// one channel of the BatchNorm forward statistics: S = sum x, Q = sum x^2 over `count` elements
__device__ __forceinline__ void bn_forward_channel(const CaeBN& bn, int c, double S, double Q, double count) {
    double mean = S / count;
    double var = Q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    double invstd = rsqrt(var + (double)bn.eps);
    float g = bn.gamma ? bn.gamma[c] : 1.f;
    float b = bn.beta ? bn.beta[c] : 0.f;
    bn.scale[c] = (float)((double)g * invstd);
    bn.shift[c] = (float)((double)b - mean * (double)g * invstd);
    bn.mean[c] = (float)mean;
    bn.invstd[c] = (float)invstd;
    if (bn.running_mean) {
        double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        double m = (double)bn.momentum;
        bn.running_mean[c] = (float)((1.0 - m) * (double)bn.running_mean[c] + m * mean);
        bn.running_var[c] = (float)((1.0 - m) * (double)bn.running_var[c] + m * unbiased);
    }
}

__device__ __forceinline__ void finalize_bn_forward(const CaeBN& bn, const double* part, int rows, double count) {
    for_each_channel_sums(part, rows, bn.C, [&](int c, double S, double Q) { bn_forward_channel(bn, c, S, Q, count); });
    if (threadIdx.x == 0 && bn.num_batches_tracked) bn.num_batches_tracked[0] += 1;
}

// BatchNorm backward sums -> dgamma, dbeta and the coefficients of dL/dy = A*dz + B*y + C
// one channel of the BatchNorm backward sums: S1 = sum dz, S2 = sum dz * xhat
__device__ __forceinline__ void bn_backward_channel(const CaeBN& bn, int c, double S1, double S2, double count) {
    double g = bn.gamma ? (double)bn.gamma[c] : 1.0;
    double invstd = (double)bn.invstd[c], mean = (double)bn.mean[c];
    double A = g * invstd;
    double B = -A * invstd * S2 / count;
    double Cc = -A * S1 / count - B * mean;
    bn.bwdA[c] = (float)A;
    bn.bwdB[c] = (float)B;
    bn.bwdC[c] = (float)Cc;
    if (bn.dgamma) bn.dgamma[c] = (float)S2;
    if (bn.dbeta) bn.dbeta[c] = (float)S1;
    // the bias of a conv that feeds a training-mode BN has an identically zero gradient
    // (sum_y dL/dy = 0); autograd returns rounding noise here.
    if (bn.dbias) bn.dbias[c] = 0.f;
}

__device__ __forceinline__ void finalize_bn_backward(const CaeBN& bn, const double* part, int rows, double count) {
    for_each_channel_sums(part, rows, bn.C, [&](int c, double S1, double S2) { bn_backward_channel(bn, c, S1, S2, count); });
}
