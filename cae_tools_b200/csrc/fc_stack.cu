// The fully connected bottleneck of the auto-encoder as ONE kernel per direction.
//   ConvAEModel : Linear(flat, fc) ReLU Linear(fc, latent) | Linear(latent, fc) ReLU Linear(fc, flat0)
//                 (reference encoder.py:52-58, decoder.py:29-35)
//   UNET        : Linear BatchNorm1d ReLU Linear ReLU | Linear BatchNorm1d ReLU Linear ReLU   (unet.py:92-100,121-129)
// At the sizes cae_tools trains with by default (batch <= 256, fc 16-128, latent 4-32) the four GEMMs are < 1 MFLOP:
// launched one by one (4 + 2 BatchNorm launches forward, 8 + 4 backward) they cost 60 us of pure launch + L2 latency per
// step.  Here a single CTA keeps every intermediate in shared memory; all sums run in a fixed order (deterministic).
// Larger bottlenecks (cae_fc_stack_supported() == 0) keep the cae_gemm chain.
#include "capi_host.h"

#define FS_NT 512
#define FS_MAX_SMEM (222 * 1024)

// Everything the kernels touch more than once is staged into shared memory with coalesced cooperative loads first
// (a first version read the weights row-per-thread from global memory: 32 L1 wavefronts per warp load and one exposed
// L2 latency per loop iteration made the single CTA slower than the eleven launches it replaced).
struct FsLayout {     // offsets in floats
    int As, t1, a1, z, t3, a3, bn, W1, W2, W3, W4, du, gA, gz, total;
};

__host__ __device__ static inline int fs_pad(int n) { return (n + 3) & ~3; }
__host__ __device__ static inline int fs_odd(int n) { return n | 1; }

__host__ __device__ static inline FsLayout fs_layout(const CaeFcStack& p, bool backward) {
    FsLayout L;
    int off = 0;
    L.As = off; off += fs_pad(p.N * p.in1);
    L.t1 = off; off += fs_pad(p.N * p.fc1);
    L.a1 = off; off += fs_pad(p.N * p.fc1);
    L.z = off; off += fs_pad(p.N * p.lat);
    L.t3 = off; off += fs_pad(p.N * p.fc2);
    L.a3 = off; off += fs_pad(p.N * p.fc2);
    L.bn = off; off += fs_pad(2 * (p.fc1 + p.fc2));
    if (!backward) {          // forward: transposed weights Wt[k][o], odd pitch (conflict-free transposing store)
        L.W1 = off; off += fs_pad(p.in1 * fs_odd(p.fc1));
        L.W2 = off; off += fs_pad(p.fc1 * fs_odd(p.lat));
        L.W3 = off; off += fs_pad(p.lat * fs_odd(p.fc2));
        L.W4 = off; off += fs_pad(p.fc2 * fs_odd(p.out4));
        L.du = L.gA = L.gz = off;
    } else {                  // backward: natural layout W[o][k]
        L.W1 = off; off += fs_pad(p.fc1 * p.in1);
        L.W2 = off; off += fs_pad(p.lat * p.fc1);
        L.W3 = off; off += fs_pad(p.fc2 * p.lat);
        L.W4 = off; off += fs_pad(p.out4 * p.fc2);
        L.du = off; off += fs_pad(p.N * p.out4);
        L.gA = off; off += fs_pad(p.N * (p.fc1 > p.fc2 ? p.fc1 : p.fc2));
        L.gz = off; off += fs_pad(p.N * p.lat);
    }
    L.total = off;
    return L;
}

__device__ __forceinline__ void fs_copy(float* dst, const float* src, int n) {
    for (int e = threadIdx.x; e < n; e += FS_NT) dst[e] = __ldg(src + e);
}

// Wt[k * OP + o] = W[o * K + k]   (global reads coalesced along k, shared stores at an odd stride)
__device__ __forceinline__ void fs_stage_Wt(float* Wt, const float* W, int O, int K) {
    const int OP = fs_odd(O);
    for (int e = threadIdx.x; e < O * K; e += FS_NT) {
        const int o = e / K, k = e - o * K;
        Wt[k * OP + o] = __ldg(W + e);
    }
}

// BatchNorm1d forward over the batch for every feature of t[N][F] (one warp per feature), training mode:
// scale/shift into shared memory + the BN scratch, running statistics updated like nn.BatchNorm1d.train()
__device__ __forceinline__ void fs_bn_train(const CaeBN& bn, const float* t, int N, int F, float* sc, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = warp; f < F; f += FS_NT / 32) {
        double s = 0.0, q = 0.0;
        for (int n = lane; n < N; n += 32) {
            const double v = (double)t[n * F + f];
            s += v;
            q += v * v;
        }
        s = warp_sum_d(s);
        q = warp_sum_d(q);
        if (lane == 0) {
            const double mean = s / N;
            double var = q / N - mean * mean;
            if (var < 0.0) var = 0.0;
            const double invstd = rsqrt(var + (double)bn.eps);
            const double g = bn.gamma ? (double)bn.gamma[f] : 1.0, b = bn.beta ? (double)bn.beta[f] : 0.0;
            sc[f] = (float)(g * invstd);
            sh[f] = (float)(b - mean * g * invstd);
            bn.scale[f] = sc[f];
            bn.shift[f] = sh[f];
            bn.mean[f] = (float)mean;
            bn.invstd[f] = (float)invstd;
            if (bn.running_mean) {
                const double unbiased = N > 1 ? var * N / (N - 1.0) : var, m = (double)bn.momentum;
                bn.running_mean[f] = (float)((1.0 - m) * (double)bn.running_mean[f] + m * mean);
                bn.running_var[f] = (float)((1.0 - m) * (double)bn.running_var[f] + m * unbiased);
            }
        }
    }
    if (threadIdx.x == 0 && bn.num_batches_tracked) bn.num_batches_tracked[0] += 1;
}

__device__ __forceinline__ void fs_load_affine(const CaeBN& bn, bool has_bn, int F, float* sc, float* sh) {
    for (int f = threadIdx.x; f < F; f += FS_NT) {
        sc[f] = has_bn ? bn.scale[f] : 1.f;
        sh[f] = has_bn ? bn.shift[f] : 0.f;
    }
}

// stage A[n][k] = relu?(k0[c] * A + k2[c]), c = k / a_hw
__device__ __forceinline__ void fs_stage_A(const CaeFcStack& p, float* As) {
    for (int e = threadIdx.x; e < p.N * p.in1; e += FS_NT) {
        const int k = e % p.in1;
        float v = __ldg(p.A + e);
        if (p.a_k0) {
            const int c = k / p.a_hw;
            v = fmaf(v, __ldg(p.a_k0 + c), __ldg(p.a_k2 + c));
        }
        if (p.a_relu) v = fmaxf(v, 0.f);
        As[e] = v;
    }
}

// out[n][o] = b[o] + sum_k in[n][k] * Wt[k][o]   (everything in shared memory; o fastest: Wt reads conflict-free,
// in[] reads broadcast; 2 samples per thread share every weight load)
__device__ __forceinline__ void fs_linear(const float* in, int N, int K, const float* Wt, const float* b, int O, float* out_s,
                                          float* out_g, bool relu) {
    const int OP = fs_odd(O), NH = (N + 1) / 2;
    for (int e = threadIdx.x; e < NH * O; e += FS_NT) {
        const int nh = e / O, o = e - nh * O;
        const int n0 = nh, n1 = nh + NH;
        const bool two = n1 < N;
        const float* x0 = in + n0 * K;
        const float* x1 = in + (two ? n1 : n0) * K;
        const float bb = b ? __ldg(b + o) : 0.f;
        float a0 = bb, a1 = bb;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float w = Wt[k * OP + o];
            a0 = fmaf(x0[k], w, a0);
            a1 = fmaf(x1[k], w, a1);
        }
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
        if (out_s) { out_s[n0 * O + o] = a0; if (two) out_s[n1 * O + o] = a1; }
        if (out_g) { out_g[n0 * O + o] = a0; if (two) out_g[n1 * O + o] = a1; }
    }
}

__global__ void __launch_bounds__(FS_NT, 1) k_fc_stack_fwd(const CaeFcStack p) {
    extern __shared__ __align__(16) float sm[];
    const FsLayout L = fs_layout(p, false);
    const bool has_bn = p.bn1.C > 0;
    float* sc1 = sm + L.bn, *sh1 = sc1 + p.fc1, *sc3 = sh1 + p.fc1, *sh3 = sc3 + p.fc2;
    fs_stage_A(p, sm + L.As);
    fs_stage_Wt(sm + L.W1, p.W1, p.fc1, p.in1);
    fs_stage_Wt(sm + L.W2, p.W2, p.lat, p.fc1);
    fs_stage_Wt(sm + L.W3, p.W3, p.fc2, p.lat);
    fs_stage_Wt(sm + L.W4, p.W4, p.out4, p.fc2);
    if (!(has_bn && p.train)) {
        fs_load_affine(p.bn1, has_bn, p.fc1, sc1, sh1);
        fs_load_affine(p.bn3, has_bn, p.fc2, sc3, sh3);
    }
    __syncthreads();
    // Linear 1 (pre-activation kept: BatchNorm / ReLU are differentiated from it)
    fs_linear(sm + L.As, p.N, p.in1, sm + L.W1, p.b1, p.fc1, sm + L.t1, p.t1, false);
    __syncthreads();
    if (has_bn && p.train) {
        fs_bn_train(p.bn1, sm + L.t1, p.N, p.fc1, sc1, sh1);
        __syncthreads();
    }
    for (int e = threadIdx.x; e < p.N * p.fc1; e += FS_NT)
        sm[L.a1 + e] = fmaxf(fmaf(sm[L.t1 + e], sc1[e % p.fc1], sh1[e % p.fc1]), 0.f);
    __syncthreads();
    fs_linear(sm + L.a1, p.N, p.fc1, sm + L.W2, p.b2, p.lat, sm + L.z, p.z, p.relu_mid != 0);
    __syncthreads();
    fs_linear(sm + L.z, p.N, p.lat, sm + L.W3, p.b3, p.fc2, sm + L.t3, p.t3, false);
    __syncthreads();
    if (has_bn && p.train) {
        fs_bn_train(p.bn3, sm + L.t3, p.N, p.fc2, sc3, sh3);
        __syncthreads();
    }
    for (int e = threadIdx.x; e < p.N * p.fc2; e += FS_NT)
        sm[L.a3 + e] = fmaxf(fmaf(sm[L.t3 + e], sc3[e % p.fc2], sh3[e % p.fc2]), 0.f);
    __syncthreads();
    fs_linear(sm + L.a3, p.N, p.fc2, sm + L.W4, p.b4, p.out4, nullptr, p.u, p.relu_mid != 0);
}

// BatchNorm1d backward for d[N][F] (already masked by the ReLU): dgamma, dbeta, d <- dL/dt  (one warp per feature)
__device__ __forceinline__ void fs_bn_bwd(const CaeBN& bn, float* d, const float* t, int N, int F) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = warp; f < F; f += FS_NT / 32) {
        const float mean = bn.mean[f], invstd = bn.invstd[f];
        double s1 = 0.0, s2 = 0.0;
        for (int n = lane; n < N; n += 32) {
            const float dv = d[n * F + f], xh = (t[n * F + f] - mean) * invstd;
            s1 += (double)dv;
            s2 += (double)(dv * xh);
        }
        s1 = warp_sum_d(s1);
        s2 = warp_sum_d(s2);
        const double g = bn.gamma ? (double)bn.gamma[f] : 1.0;
        const double A = g * (double)invstd, B = -A * (double)invstd * s2 / N, Cc = -A * s1 / N - B * (double)mean;
        for (int n = lane; n < N; n += 32)
            d[n * F + f] = (float)A * d[n * F + f] + (float)B * t[n * F + f] + (float)Cc;
        if (lane == 0) {
            if (bn.dgamma) bn.dgamma[f] = (float)s2;
            if (bn.dbeta) bn.dbeta[f] = (float)s1;
        }
    }
}

// dW[o][k] = sum_n d[n][o] * x[n][k] ; db[o] = sum_n d[n][o]    (d, x in shared memory; fixed order over n)
__device__ __forceinline__ void fs_wgrad(const float* d, const float* x, int N, int O, int K, float* dW, float* db) {
    for (int e = threadIdx.x; e < O * K; e += FS_NT) {
        const int o = e / K, k = e - o * K;
        float a0 = 0.f, a1 = 0.f;
        int n = 0;
        for (; n + 1 < N; n += 2) {
            a0 = fmaf(d[n * O + o], x[n * K + k], a0);
            a1 = fmaf(d[(n + 1) * O + o], x[(n + 1) * K + k], a1);
        }
        if (n < N) a0 = fmaf(d[n * O + o], x[n * K + k], a0);
        dW[e] = a0 + a1;
    }
    if (db)
        for (int o = threadIdx.x; o < O; o += FS_NT) {
            float acc = 0.f;
            for (int n = 0; n < N; ++n) acc += d[n * O + o];
            db[o] = acc;
        }
}

// dx[n][k] = sum_o d[n][o] * W[o][k]  (W natural layout in shared memory; optionally masked by act[n][k] > 0)
__device__ __forceinline__ void fs_dgrad(const float* d, int N, int O, const float* W, int K, const float* act, float* dx_s,
                                         float* dx_g) {
    for (int e = threadIdx.x; e < N * K; e += FS_NT) {
        const int n = e / K, k = e - n * K;
        const float* dn = d + n * O;
        float a0 = 0.f, a1 = 0.f;
        int o = 0;
        for (; o + 1 < O; o += 2) {
            a0 = fmaf(dn[o], W[o * K + k], a0);
            a1 = fmaf(dn[o + 1], W[(o + 1) * K + k], a1);
        }
        if (o < O) a0 = fmaf(dn[o], W[o * K + k], a0);
        float acc = a0 + a1;
        if (act && !(act[e] > 0.f)) acc = 0.f;
        if (dx_s) dx_s[e] = acc;
        if (dx_g) dx_g[e] = acc;
    }
}

__global__ void __launch_bounds__(FS_NT, 1) k_fc_stack_bwd(const CaeFcStack p) {
    extern __shared__ __align__(16) float sm[];
    const FsLayout L = fs_layout(p, true);
    const bool has_bn = p.bn1.C > 0;
    float* sc1 = sm + L.bn, *sh1 = sc1 + p.fc1, *sc3 = sh1 + p.fc1, *sh3 = sc3 + p.fc2;
    float* gA = sm + L.gA;                                        // [N][max(fc1, fc2)] gradient wrt t3, then t1
    float* gz = sm + L.gz;                                        // [N][lat]
    float* du = sm + L.du;                                        // [N][out4]
    fs_stage_A(p, sm + L.As);
    fs_load_affine(p.bn1, has_bn, p.fc1, sc1, sh1);
    fs_load_affine(p.bn3, has_bn, p.fc2, sc3, sh3);
    fs_copy(sm + L.t1, p.t1, p.N * p.fc1);
    fs_copy(sm + L.t3, p.t3, p.N * p.fc2);
    fs_copy(sm + L.z, p.z, p.N * p.lat);
    fs_copy(du, p.du, p.N * p.out4);
    fs_copy(sm + L.W1, p.W1, p.fc1 * p.in1);
    fs_copy(sm + L.W2, p.W2, p.lat * p.fc1);
    fs_copy(sm + L.W3, p.W3, p.fc2 * p.lat);
    fs_copy(sm + L.W4, p.W4, p.out4 * p.fc2);
    __syncthreads();
    for (int e = threadIdx.x; e < p.N * p.fc1; e += FS_NT)
        sm[L.a1 + e] = fmaxf(fmaf(sm[L.t1 + e], sc1[e % p.fc1], sh1[e % p.fc1]), 0.f);
    for (int e = threadIdx.x; e < p.N * p.fc2; e += FS_NT)
        sm[L.a3 + e] = fmaxf(fmaf(sm[L.t3 + e], sc3[e % p.fc2], sh3[e % p.fc2]), 0.f);
    __syncthreads();
    // ---- Linear 4: u = act(a3 W4^T + b4); du arrives already masked by that activation
    fs_wgrad(du, sm + L.a3, p.N, p.out4, p.fc2, p.dW4, p.db4);
    fs_dgrad(du, p.N, p.out4, sm + L.W4, p.fc2, sm + L.a3, gA, nullptr);          // gA = dL/d a3 masked by a3 > 0
    __syncthreads();
    if (has_bn) {
        fs_bn_bwd(p.bn3, gA, sm + L.t3, p.N, p.fc2);                               // gA = dL/dt3
        __syncthreads();
    }
    // ---- Linear 3: t3 = z W3^T + b3
    fs_wgrad(gA, sm + L.z, p.N, p.fc2, p.lat, p.dW3, has_bn ? nullptr : p.db3);
    fs_dgrad(gA, p.N, p.fc2, sm + L.W3, p.lat, p.relu_mid ? sm + L.z : nullptr, gz, nullptr);
    __syncthreads();
    // ---- Linear 2: z = act(a1 W2^T + b2)
    fs_wgrad(gz, sm + L.a1, p.N, p.lat, p.fc1, p.dW2, p.db2);
    fs_dgrad(gz, p.N, p.lat, sm + L.W2, p.fc1, sm + L.a1, gA, nullptr);           // (dt3 in gA fully consumed above)
    __syncthreads();
    if (has_bn) {
        fs_bn_bwd(p.bn1, gA, sm + L.t1, p.N, p.fc1);                               // gA = dL/dt1
        __syncthreads();
    }
    // ---- Linear 1: t1 = A W1^T + b1
    fs_wgrad(gA, sm + L.As, p.N, p.fc1, p.in1, p.dW1, has_bn ? nullptr : p.db1);
    fs_dgrad(gA, p.N, p.fc1, sm + L.W1, p.in1, nullptr, nullptr, p.dA);
}

// =====================================================================================================
static int fs_check(const CaeFcStack* p, bool backward, size_t* smem) {
    CAE_REQUIRE(p && p->A && p->W1 && p->W2 && p->W3 && p->W4 && p->t1 && p->z && p->t3 && p->u, "fc_stack: null argument");
    CAE_REQUIRE(p->N > 0 && p->in1 > 0 && p->fc1 > 0 && p->lat > 0 && p->fc2 > 0 && p->out4 > 0, "fc_stack: empty dimension");
    CAE_REQUIRE((p->bn1.C > 0) == (p->bn3.C > 0), "fc_stack: BatchNorm after both first Linear layers or after none");
    if (p->bn1.C > 0) {
        CAE_REQUIRE(p->bn1.C == p->fc1 && p->bn3.C == p->fc2 && p->bn1.scale && p->bn1.shift && p->bn1.mean && p->bn1.invstd &&
                        p->bn3.scale && p->bn3.shift && p->bn3.mean && p->bn3.invstd, "fc_stack: BN block incomplete");
    }
    CAE_REQUIRE(!p->a_k0 || (p->a_k2 && p->a_hw > 0), "fc_stack: on-load affine needs k0, k2 and hw");
    const FsLayout L = fs_layout(*p, backward);
    *smem = (size_t)L.total * 4;
    if (*smem > FS_MAX_SMEM) {
        cae_set_error("fc_stack: %zu KB of shared memory needed (> %d): use the cae_gemm chain", *smem / 1024, FS_MAX_SMEM / 1024);
        return CAE_EUNSUPPORTED;
    }
    if (backward) {
        CAE_REQUIRE(p->du && p->dW1 && p->dW2 && p->dW3 && p->dW4 && p->dA, "fc_stack_bwd: null gradient pointer");
    }
    return CAE_OK;
}

extern "C" int cae_fc_stack_supported(int N, int in1, int fc1, int lat, int fc2, int out4) {
    CaeFcStack p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.in1 = in1; p.fc1 = fc1; p.lat = lat; p.fc2 = fc2; p.out4 = out4;
    if (N < 1 || N > 512 || fc1 > 256 || fc2 > 256 || lat > 256 || out4 > 8192 || in1 > 8192) return 0;
    return (size_t)fs_layout(p, true).total * 4 <= FS_MAX_SMEM ? 1 : 0;
}

extern "C" int cae_fc_stack_fwd(const CaeFcStack* p, void* stream) {
    size_t smem;
    int rc = fs_check(p, false, &smem);
    if (rc) return rc;
    ensure_smem_limit(k_fc_stack_fwd, FS_MAX_SMEM);
    k_fc_stack_fwd<<<1, FS_NT, smem, (cudaStream_t)stream>>>(*p);
    return cae_check_launch("cae_fc_stack_fwd");
}

extern "C" int cae_fc_stack_bwd(const CaeFcStack* p, void* stream) {
    size_t smem;
    int rc = fs_check(p, true, &smem);
    if (rc) return rc;
    ensure_smem_limit(k_fc_stack_bwd, FS_MAX_SMEM);
    k_fc_stack_bwd<<<1, FS_NT, smem, (cudaStream_t)stream>>>(*p);
    return cae_check_launch("cae_fc_stack_bwd");
}
