// Small fp32 GEMM for the fully-connected bottleneck, eval-mode BatchNorm prepare, MSE, fused Adam.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------
// C[m,n] = epi( sum_k A(m,k) B(k,n) ) with arbitrary strides; 32x32 tile, 2x2 per thread.
// ---------------------------------------------------------------------------------------
#define GT 32
static __global__ void __launch_bounds__(CAE_NT) k_gemm(const CaeGemm g) {
    __shared__ float As[GT][GT + 1];  // [m][k]
    __shared__ float Bs[GT][GT + 1];  // [k][n]
    __shared__ float rs[GT];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 2x2
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const bool a_kfast = (g.sAk <= g.sAm);
    const bool b_nfast = (g.sBn <= g.sBk);
    const bool do_rowsum = (g.rowsum_A != nullptr) && blockIdx.x == 0;
    if (tid < GT) rs[tid] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += GT) {
        // stage A tile
#pragma unroll
        for (int i = 0; i < (GT * GT) / CAE_NT; ++i) {
            int idx = tid + i * CAE_NT;
            int mm = a_kfast ? idx / GT : idx % GT;
            int kk = a_kfast ? idx % GT : idx / GT;
            int m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < g.M && k < g.K) {
                v = __ldg(g.A + (long long)m * g.sAm + (long long)k * g.sAk);
                if (g.a_k0) {
                    int c = k / g.a_hw;
                    v = fmaf(v, __ldg(g.a_k0 + c), __ldg(g.a_k2 + c));
                }
                if (g.a_relu) v = fmaxf(v, 0.f);
            }
            As[mm][kk] = v;
        }
#pragma unroll
        for (int i = 0; i < (GT * GT) / CAE_NT; ++i) {
            int idx = tid + i * CAE_NT;
            int kk = b_nfast ? idx / GT : idx % GT;
            int nn = b_nfast ? idx % GT : idx / GT;
            int k = k0 + kk, n = n0 + nn;
            float v = 0.f;
            if (k < g.K && n < g.N) {
                v = __ldg(g.B + (long long)k * g.sBk + (long long)n * g.sBn);
                if (g.b_k0) {
                    int c = n / g.b_hw;
                    v = fmaf(v, __ldg(g.b_k0 + c), __ldg(g.b_k2 + c));
                }
                if (g.b_relu) v = fmaxf(v, 0.f);
            }
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GT; ++kk) {
            float a0 = As[ty * 2][kk], a1 = As[ty * 2 + 1][kk];
            float b0 = Bs[kk][tx * 2], b1 = Bs[kk][tx * 2 + 1];
            acc[0][0] = fmaf(a0, b0, acc[0][0]);
            acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]);
            acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        if (do_rowsum && tid < GT) {
            float s = rs[tid];
            for (int kk = 0; kk < GT; ++kk) s += As[tid][kk];
            rs[tid] = s;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
            if (m < g.M && n < g.N) {
                float v = acc[i][j];
                if (g.bias) v += __ldg(g.bias + n);
                if (g.relu_out) v = fmaxf(v, 0.f);
                long long off = (long long)m * g.sCm + (long long)n * g.sCn;
                if (g.mask) v = __ldg(g.mask + off) > 0.f ? v : 0.f;
                g.C[off] = v;
            }
        }
    if (do_rowsum && tid < GT && m0 + tid < g.M) g.rowsum_A[m0 + tid] = rs[tid];
}

// Skinny problems (few outputs, long K): one thread per output element, K unrolled so the (L1/L2 resident)
// operand loads overlap.  No on-load transforms, no row sums - those stay with the tiled kernel.
static __global__ void __launch_bounds__(CAE_NT) k_gemm_skinny(const CaeGemm g) {
    const int idx = blockIdx.x * CAE_NT + threadIdx.x;
    if (idx >= g.M * g.N) return;
    const int m = idx / g.N, n = idx - m * g.N;
    const float* ap = g.A + (long long)m * g.sAm;
    const float* bp = g.B + (long long)n * g.sBn;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int k = 0;
    for (; k + 3 < g.K; k += 4) {
        float a0 = __ldg(ap + (long long)k * g.sAk), a1 = __ldg(ap + (long long)(k + 1) * g.sAk);
        float a2 = __ldg(ap + (long long)(k + 2) * g.sAk), a3 = __ldg(ap + (long long)(k + 3) * g.sAk);
        float b0 = __ldg(bp + (long long)k * g.sBk), b1 = __ldg(bp + (long long)(k + 1) * g.sBk);
        float b2 = __ldg(bp + (long long)(k + 2) * g.sBk), b3 = __ldg(bp + (long long)(k + 3) * g.sBk);
        acc0 = fmaf(a0, b0, acc0); acc1 = fmaf(a1, b1, acc1); acc2 = fmaf(a2, b2, acc2); acc3 = fmaf(a3, b3, acc3);
    }
    for (; k < g.K; ++k) acc0 = fmaf(__ldg(ap + (long long)k * g.sAk), __ldg(bp + (long long)k * g.sBk), acc0);
    float v = (acc0 + acc1) + (acc2 + acc3);
    if (g.bias) v += __ldg(g.bias + n);
    if (g.relu_out) v = fmaxf(v, 0.f);
    const long long off = (long long)m * g.sCm + (long long)n * g.sCn;
    if (g.mask) v = __ldg(g.mask + off) > 0.f ? v : 0.f;
    g.C[off] = v;
}

// Few outputs, long reduction (the ConvAE bottleneck's input gradient: M x N = 64 x 16 outputs, K = 576): one WARP per
// output element, lanes over k, butterfly sum.  The thread-per-output kernel above walks K serially (37 us for this
// shape, two dependent loads per step); here every lane does K / 32 steps.
static __global__ void __launch_bounds__(CAE_NT) k_gemm_warpk(const CaeGemm g) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * CAE_NWARP + (threadIdx.x >> 5);
    if (idx >= g.M * g.N) return;
    const int m = idx / g.N, n = idx - m * g.N;
    const float* ap = g.A + (long long)m * g.sAm;
    const float* bp = g.B + (long long)n * g.sBn;
    float acc0 = 0.f, acc1 = 0.f;
    int k = lane;
    for (; k + 32 < g.K; k += 64) {
        const float a0 = __ldg(ap + (long long)k * g.sAk), a1 = __ldg(ap + (long long)(k + 32) * g.sAk);
        const float b0 = __ldg(bp + (long long)k * g.sBk), b1 = __ldg(bp + (long long)(k + 32) * g.sBk);
        acc0 = fmaf(a0, b0, acc0);
        acc1 = fmaf(a1, b1, acc1);
    }
    for (; k < g.K; k += 32) acc0 = fmaf(__ldg(ap + (long long)k * g.sAk), __ldg(bp + (long long)k * g.sBk), acc0);
    float v = warp_sum(acc0 + acc1);
    if (lane == 0) {
        if (g.bias) v += __ldg(g.bias + n);
        if (g.relu_out) v = fmaxf(v, 0.f);
        const long long off = (long long)m * g.sCm + (long long)n * g.sCn;
        if (g.mask) v = __ldg(g.mask + off) > 0.f ? v : 0.f;
        g.C[off] = v;
    }
}

// ---------------------------------------------------------------------------------------
// eval-mode BatchNorm: scale/shift from the running statistics; one CTA per layer
// ---------------------------------------------------------------------------------------
static __global__ void k_bn_eval_prepare(const CaeBN* table, int count) {
    const CaeBN bn = table[blockIdx.x];
    for (int c = threadIdx.x; c < bn.C; c += blockDim.x) {
        float invstd = 1.f / sqrtf(bn.running_var[c] + bn.eps);
        float g = bn.gamma ? bn.gamma[c] : 1.f;
        float b = bn.beta ? bn.beta[c] : 0.f;
        float sc = g * invstd;
        bn.scale[c] = sc;
        bn.shift[c] = b - bn.running_mean[c] * sc;
    }
}

// ---------------------------------------------------------------------------------------
// MSE over flat arrays, deterministic
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_mse(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                double* partials, unsigned int* ticket, float* loss_out,
                                                const int* cursor) {
    float s = 0.f;
    double sd = 0.0;
    int it = 0;
    for (long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CAE_NT) {
        float d = __ldg(a + i) - __ldg(b + i);
        s = fmaf(d, d, s);
        if (++it == 64) {  // bound fp32 accumulation length
            sd += (double)s;
            s = 0.f;
            it = 0;
        }
    }
    sd += (double)s;
    __shared__ double red[CAE_NWARP];
    sd = warp_sum_d(sd);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < CAE_NWARP; ++w) t += red[w];
        partials[blockIdx.x] = t;
    }
    if (cae_last_block(ticket)) {
        if (threadIdx.x < 32) {
            double t = 0.0;
            for (int r = threadIdx.x; r < (int)gridDim.x; r += 32) t += __ldcg(partials + r);
            t = warp_sum_d(t);
            if (threadIdx.x == 0) loss_out[cursor ? __ldg(cursor) : 0] = (float)(t / (double)n);
        }
    }
}

// ---------------------------------------------------------------------------------------
// fused multi-tensor Adam / AdamW over the flat parameter arena
// (update order follows torch.optim.adam._single_tensor_adam)
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                 float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                 float eps, float wd, int decoupled, float gscale,
                                                 const int* step_count) {
    __shared__ float s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        double t = (double)(__ldg(step_count) + 1);
        double bc1 = 1.0 - pow((double)beta1, t);
        double bc2 = 1.0 - pow((double)beta2, t);
        s_step_size = (float)((double)lr / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    for (long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CAE_NT) {
        float pi = p[i], gi = g[i] * gscale, mi = m[i], vi = v[i];
        if (decoupled) pi *= (1.f - lr * wd);
        else gi = fmaf(wd, pi, gi);
        mi = fmaf(gi - mi, 1.f - beta1, mi);            // exp_avg.lerp_(grad, 1-beta1)
        vi = fmaf((1.f - beta2) * gi, gi, vi * beta2);  // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
    }
}

// Adam + the step bookkeeping in ONE launch: the CTA that finishes last (ticket) advances the step counter and the batch
// cursor - no CTA can still be reading step_count then (the separate one-thread k_step_advance launch cost ~6 us of a
// ~220 us training step).
static __global__ void __launch_bounds__(CAE_NT) k_adam_advance(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                         float eps, float wd, int decoupled, float gscale, int* step_count,
                                                         int* cursor, int n_batches, unsigned int* ticket) {
    __shared__ float s_step_size, s_bc2_sqrt;
    // the first element's four loads are issued BEFORE thread 0's two float64 pow() calls (~1.5 us on one thread) so that
    // their latency hides behind them (the arena of the bench workload is one element per thread)
    long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x;
    float pi = 0.f, gi = 0.f, mi = 0.f, vi = 0.f;
    if (i < n) { pi = p[i]; gi = g[i]; mi = m[i]; vi = v[i]; }
    if (threadIdx.x == 0) {
        double t = (double)(*reinterpret_cast<volatile int*>(step_count) + 1);
        double bc1 = 1.0 - pow((double)beta1, t);
        double bc2 = 1.0 - pow((double)beta2, t);
        s_step_size = (float)((double)lr / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    while (i < n) {
        gi *= gscale;
        if (decoupled) pi *= (1.f - lr * wd);
        else gi = fmaf(wd, pi, gi);
        mi = fmaf(gi - mi, 1.f - beta1, mi);
        vi = fmaf((1.f - beta2) * gi, gi, vi * beta2);
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        i += (long long)gridDim.x * CAE_NT;
        if (i < n) { pi = p[i]; gi = g[i]; mi = m[i]; vi = v[i]; }
    }
    if (cae_last_block(ticket) && threadIdx.x == 0) {
        step_count[0] += 1;
        if (cursor) {
            int c = cursor[0] + 1;
            cursor[0] = (c >= n_batches) ? 0 : c;
        }
    }
}

static __global__ void k_step_advance(int* step_count, int* cursor, int n_batches) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (step_count) step_count[0] += 1;
        if (cursor) {
            int c = cursor[0] + 1;
            cursor[0] = (c >= n_batches) ? 0 : c;
        }
    }
}

// ---------------------------------------------------------------------------------------
// variational bottleneck: z = mu + eps * exp(logvar/2) ; KL = -1/2 sum(1 + logvar - mu^2 - exp(logvar)) / N
// (single CTA: the latent block is N x L with L <= a few thousand)
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CAE_NT) k_vae_reparam_fwd(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                            const float* __restrict__ eps, long long eps_stride,
                                                            const int* cursor, float* __restrict__ z, int n, int sample,
                                                            float inv_n, float* kl_out) {
    const float* e = eps + (cursor ? (long long)__ldg(cursor) * eps_stride : 0ll);
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += CAE_NT) {
        float m = mu[i], lv = logvar[i];
        float sd = expf(0.5f * lv);
        if (sample) z[i] = fmaf(__ldg(e + i), sd, m);
        else z[i] = m;
        acc += (double)(1.f + lv - m * m - sd * sd);
    }
    __shared__ double red[CAE_NWARP];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && kl_out) {
        double t = 0.0;
        for (int w = 0; w < CAE_NWARP; ++w) t += red[w];
        kl_out[cursor ? __ldg(cursor) : 0] = (float)(-0.5 * t * (double)inv_n);
    }
}

// dmu = dz + kl_w * mu ; dlogvar = dz * eps * exp(logvar/2) / 2 + kl_w * (exp(logvar) - 1) / 2,  kl_w = lambda_kl / N
static __global__ void __launch_bounds__(CAE_NT) k_vae_reparam_bwd(const float* __restrict__ dz, const float* __restrict__ mu,
                                                            const float* __restrict__ logvar, const float* __restrict__ eps,
                                                            long long eps_stride, const int* cursor,
                                                            float* __restrict__ dmu, float* __restrict__ dlogvar, int n,
                                                            float kl_w) {
    const float* e = eps + (cursor ? (long long)__ldg(cursor) * eps_stride : 0ll);
    for (int i = blockIdx.x * CAE_NT + threadIdx.x; i < n; i += gridDim.x * CAE_NT) {
        float m = mu[i], lv = logvar[i], g = dz[i];
        float sd = expf(0.5f * lv);
        dmu[i] = fmaf(kl_w, m, g);
        dlogvar[i] = 0.5f * g * __ldg(e + i) * sd + 0.5f * kl_w * (sd * sd - 1.f);
    }
}

// out[i] = a[i] + b[i]   (gradient fan-in of the two latent heads)
static __global__ void __launch_bounds__(CAE_NT) k_add2(const float* __restrict__ a, const float* __restrict__ b,
                                                 float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CAE_NT)
        out[i] = a[i] + b[i];
}

// ---------------------------------------------------------------------------------------
// standard-normal fill: counter-based (seed, step, index) -> splitmix64 -> Box-Muller.  Stateless, so a
// captured graph draws fresh noise on every replay (the step counter lives on the device).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cae_mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static __global__ void __launch_bounds__(CAE_NT) k_randn(float* __restrict__ out, long long n, unsigned long long seed,
                                                  const int* step_count) {
    const unsigned long long step = step_count ? (unsigned long long)__ldg(step_count) : 0ull;
    const unsigned long long key = cae_mix64(seed ^ cae_mix64(step + 0x51ED270B5ull));
    for (long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CAE_NT) {
        unsigned long long r = cae_mix64(key + 2ull * (unsigned long long)i);
        unsigned int a = (unsigned int)(r >> 32), b = (unsigned int)r;
        float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
        float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);            // [0, 1)
        out[i] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
}
