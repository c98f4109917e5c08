// Data ingest on the device: the steps right before / after the hot path that the reference runs on
// the host per sample (DSDataset.__init__ min/max + NaN scan: ds_dataset.py:49-75; normalise_input / normalise_output +
// default collate of shuffled items: ds_dataset.py:99-113,137-159).  All three are single HBM-bound passes:
//   cae_minmax            : min, max and NaN count of a raw fp32 array (two-stage, fixed-order; min / max are
//                           order-independent, the count is an integer)
//   cae_normalise_gather  : dst[i] = (src[order[i]] - lo) / (hi - lo)  - min-max normalisation + batch assembly in the
//                           shuffled order in one pass, written straight into a channel slice of the batch tensor
// Arithmetic is the reference's, in fp32 with IEEE division: bit-identical to numpy on the same fp32 inputs.
#include "capi_host.h"

namespace {

__global__ void __launch_bounds__(CAE_NT) k_minmax(const float* __restrict__ x, long long n, float* __restrict__ part,
                                                   unsigned int* ticket, float* __restrict__ out3) {
    float lo = INFINITY, hi = -INFINITY;
    unsigned int nan = 0;
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * CAE_NT) {
        const float4 v = __ldg(x4 + i);
        const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (a[j] != a[j]) ++nan;
            else { lo = fminf(lo, a[j]); hi = fmaxf(hi, a[j]); }
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CAE_NT) {
        const float a = __ldg(x + i);
        if (a != a) ++nan;
        else { lo = fminf(lo, a); hi = fmaxf(hi, a); }
    }
    __shared__ float s_lo[CAE_NWARP], s_hi[CAE_NWARP];
    __shared__ unsigned int s_nan[CAE_NWARP];
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan += __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; s_nan[threadIdx.x >> 5] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < CAE_NWARP; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); nan += s_nan[w]; }
        part[blockIdx.x * 3 + 0] = lo;
        part[blockIdx.x * 3 + 1] = hi;
        part[blockIdx.x * 3 + 2] = __uint_as_float(nan);
    }
    if (cae_last_block(ticket) && threadIdx.x == 0) {
        float l = INFINITY, h = -INFINITY;
        double c = 0.0;
        for (unsigned int r = 0; r < gridDim.x; ++r) {
            l = fminf(l, __ldcg(part + r * 3));
            h = fmaxf(h, __ldcg(part + r * 3 + 1));
            c += (double)__float_as_uint(__ldcg(part + r * 3 + 2));
        }
        out3[0] = l;
        out3[1] = h;
        out3[2] = (float)c;
    }
}

// one sample of `elems` contiguous floats per CTA-row; grid.y = output samples
__global__ void __launch_bounds__(CAE_NT) k_normalise_gather(const float* __restrict__ src, long long elems, const int* __restrict__ order,
                                                             int n_out, float lo, float range, int normalise,
                                                             float* __restrict__ dst, long long dst_stride) {
    for (int i = blockIdx.y; i < n_out; i += gridDim.y) {
        const long long s = order ? (long long)__ldg(order + i) : (long long)i;
        const float* sp = src + s * elems;
        float* dp = dst + (long long)i * dst_stride;
        for (long long e = (long long)blockIdx.x * CAE_NT + threadIdx.x; e < elems; e += (long long)gridDim.x * CAE_NT) {
            const float v = __ldg(sp + e);
            dp[e] = !normalise ? v : (range == 0.f ? 0.f : __fdiv_rn(__fsub_rn(v, lo), range));
        }
    }
}

}  // namespace

extern "C" long long cae_minmax_partials_len(void) { return (long long)CAE_MAX_GRID_X * 3; }

extern "C" int cae_minmax(const float* x, long long n, float* partials, unsigned int* ticket, float* out3, void* stream) {
    CAE_REQUIRE(x && partials && ticket && out3 && n > 0, "minmax: bad argument");
    CAE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "minmax: input must be 16-byte aligned");
    const int grid = (int)min((long long)CAE_MAX_GRID_X, (n / 4 + CAE_NT - 1) / CAE_NT + 1);
    k_minmax<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(x, n, partials, ticket, out3);
    return cae_check_launch("cae_minmax");
}

extern "C" int cae_normalise_gather(const float* src, long long sample_elems, const int* order, int n_out, float lo, float hi,
                                    int normalise, float* dst, long long dst_sample_stride, void* stream) {
    CAE_REQUIRE(src && dst && sample_elems > 0 && n_out > 0 && dst_sample_stride >= sample_elems, "normalise_gather: bad argument");
    const float range = hi - lo;      // fp32 difference of the fp32-rounded bounds, like numpy's weak-scalar arithmetic
    dim3 grid((unsigned)min((long long)64, (sample_elems + CAE_NT - 1) / CAE_NT), (unsigned)min(n_out, 16384));
    k_normalise_gather<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(src, sample_elems, order, n_out, lo, range, normalise, dst,
                                                                 dst_sample_stride);
    return cae_check_launch("cae_normalise_gather");
}

// ---- post-training metrics on the device (reference model_metric.py:19-71, base_model.py:116-125) ----------
// One row of eight float64 sums per case over its kept pixels (mask > 0 or no mask): count, sum a, sum e, sum a*a, sum e*e,
// sum a*e, sum |a - e|, sum (a - e)^2 with a = actual (raw fp32) and e = lo + yhat * scale evaluated in float64 exactly as the
// reference de-normalises its float32 predictions (ds_dataset.py:122-125).  The host turns the rows into mse / rmse / mae and
// the mean per-case Pearson correlation.  One CTA per case; per-thread float64 partial sums, fixed-order CTA reduction.
__global__ void __launch_bounds__(CAE_NT) k_case_metrics(const float* __restrict__ yhat, const float* __restrict__ actual,
                                                         const float* __restrict__ mask, long long per_case, long long mask_per_case,
                                                         double lo, double scale, double* __restrict__ out) {
    __shared__ double red[CAE_NWARP][8];
    const long long n = blockIdx.x;
    const float* y = yhat + n * per_case;
    const float* a = actual + n * per_case;
    const float* m = mask ? mask + n * mask_per_case : nullptr;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = threadIdx.x; i < per_case; i += CAE_NT) {
        if (m && !(m[i % mask_per_case] != 0.f)) continue;
        const double av = (double)a[i], ev = lo + (double)y[i] * scale, d = av - ev;
        s[0] += 1.0; s[1] += av; s[2] += ev; s[3] += av * av; s[4] += ev * ev; s[5] += av * ev; s[6] += fabs(d); s[7] += d * d;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double t = warp_sum_d(s[k]);
        if (lane == 0) red[warp][k] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < CAE_NWARP; ++w) t += red[w][threadIdx.x];
        out[n * 8 + threadIdx.x] = t;
    }
}

extern "C" int cae_case_metrics(const float* yhat, const float* actual, const float* mask, int n_cases, long long per_case,
                                long long mask_per_case, double lo, double scale, double* out, void* stream) {
    CAE_REQUIRE(yhat && actual && out && n_cases > 0 && per_case > 0, "case_metrics: bad argument");
    CAE_REQUIRE(!mask || (mask_per_case > 0 && per_case % mask_per_case == 0), "case_metrics: the mask must tile a case");
    k_case_metrics<<<n_cases, CAE_NT, 0, (cudaStream_t)stream>>>(yhat, actual, mask, per_case, mask ? mask_per_case : per_case, lo, scale, out);
    return cae_check_launch("cae_case_metrics");
}
