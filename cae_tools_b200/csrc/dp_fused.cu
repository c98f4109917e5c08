// Data-parallel exchange fused into the optimiser: ONE kernel per rank does the cross-GPU gradient all-reduce (peer reads
// over NVLink / NVSwitch of every rank's flat gradient arena, summed in rank order - identical bits on every rank) and the
// Adam / AdamW update + step bookkeeping.  For the small arenas of the 16x16 -> 256x256 configs (141-157 KB) an NCCL
// all-reduce is pure latency (20 us at 2 GPUs, 39 us at 8, on a 0.22 ms step) sitting between the backward pass and the
// optimiser; here the exchange is two flag round trips and W-1 peer reads of the arena inside the optimiser launch, and the
// whole step is one CUDA graph again.  Arenas beyond a few MB stay on NCCL (every rank reads W-1 full arenas here).
//
// Protocol (flags live in a symmetric allocation: flags[r] is rank r's array, peer-mapped everywhere; epoch e is a private
// device counter that grows by one per launch on every rank, so nothing is ever reset):
//   1. CTA 0: st.release.sys  flags[peer][rank] = e            "my gradients are complete"  (to every peer, self included)
//   2. every CTA: spin until  flags[self][r] >= e  for all r    "everyone's gradients are complete"
//   3. g[i] = sum_r grads[r][i] (ld.relaxed.sys, rank order), Adam update of the local parameter replica
//   4. last CTA (ticket): st.release.sys  flags[peer][W_MAX + rank] = e   "I have read your gradients"; advance epoch,
//      step counter and batch cursor.  Nobody waits here: the matching wait is k_dp_wait_done, the FIRST launch of the
//      next step (long before that step's backward overwrites the gradients) - by then the flags have been set for
//      ~100 us, so the wait is free, yet no rank can overwrite gradients a slow peer is still reading.
// Every rank runs these kernels on its own GPU (one process per GPU), so the spins always have a live partner.
// Measured on 2 x B200 (tools/dp_fused_probe.py, 154 KB arena): Adam alone 4.8 us, NCCL all-reduce alone 18.5 us; first
// version of this kernel (second handshake at its end, scalar peer loads, explicit membar.sys) 18.5 us.
#include "capi_host.h"

#define CAE_DP_MAX_WORLD 8

struct DpArgs {
    CaeDpPeers peers;
    float* p;
    float* m;
    float* v;
    long long n;
    float lr, beta1, beta2, eps, wd, gscale;
    int decoupled;
    int* step_count;
    int* cursor;
    int n_batches;
    unsigned int* epoch;
    unsigned int* ticket;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(CAE_NT) k_adam_allreduce(const DpArgs a) {
    __shared__ float s_step_size, s_bc2_sqrt;
    __shared__ unsigned int s_epoch;
    const int W = a.peers.world, rank = a.peers.rank;
    if (threadIdx.x == 0) {
        const double t = (double)(*reinterpret_cast<volatile int*>(a.step_count) + 1);
        s_step_size = (float)((double)a.lr / (1.0 - pow((double)a.beta1, t)));
        s_bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, t));
        s_epoch = *reinterpret_cast<volatile unsigned int*>(a.epoch) + 1u;
    }
    __syncthreads();
    const unsigned int e = s_epoch;
    // 1. publish "ready" (the gradients were written by earlier kernels of this stream: complete at this launch's start)
    if (blockIdx.x == 0 && (int)threadIdx.x < W) st_release_sys(a.peers.flags[threadIdx.x] + rank, e);
    // 2. wait for every rank
    if ((int)threadIdx.x < W) {
        const unsigned int* f = a.peers.flags[rank] + threadIdx.x;
        while ((int)(ld_acquire_sys(f) - e) < 0) {}
    }
    __syncthreads();
    // 3. all-reduce + Adam
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    auto update = [&](float gi, float& pi, float& mi, float& vi) {
        gi *= a.gscale;
        if (a.decoupled) pi *= (1.f - a.lr * a.wd);
        else gi = fmaf(a.wd, pi, gi);
        mi = fmaf(gi - mi, 1.f - a.beta1, mi);
        vi = fmaf((1.f - a.beta2) * gi, gi, vi * a.beta2);
        const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
        pi -= step_size * (mi / denom);
    };
    // (the arena is a multiple of 4 floats and 16-byte aligned: whole float4 groups; every peer load of a thread is issued
    //  before the first add, so a thread pays one NVLink round trip)
    const long long n4 = a.n >> 2;
    for (long long i = (long long)blockIdx.x * CAE_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * CAE_NT) {
        float4 g[CAE_DP_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < CAE_DP_MAX_WORLD; ++r)
            if (r < W) g[r] = ld_relaxed_sys4(a.peers.grads[r] + 4 * i);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < CAE_DP_MAX_WORLD; ++r)
            if (r < W) { s.x += g[r].x; s.y += g[r].y; s.z += g[r].z; s.w += g[r].w; }
        float4 p4 = reinterpret_cast<float4*>(a.p)[i], m4 = reinterpret_cast<float4*>(a.m)[i], v4 = reinterpret_cast<float4*>(a.v)[i];
        update(s.x, p4.x, m4.x, v4.x); update(s.y, p4.y, m4.y, v4.y); update(s.z, p4.z, m4.z, v4.z); update(s.w, p4.w, m4.w, v4.w);
        reinterpret_cast<float4*>(a.p)[i] = p4;
        reinterpret_cast<float4*>(a.m)[i] = m4;
        reinterpret_cast<float4*>(a.v)[i] = v4;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * CAE_NT + threadIdx.x; i < a.n; i += (long long)gridDim.x * CAE_NT) {
        float gi = 0.f;
        for (int r = 0; r < W; ++r) gi += ld_relaxed_sys(a.peers.grads[r] + i);
        float pi = a.p[i], mi = a.m[i], vi = a.v[i];
        update(gi, pi, mi, vi);
        a.p[i] = pi; a.m[i] = mi; a.v[i] = vi;
    }
    // 4. "done reading" flags by the last CTA of this rank (nobody waits here: k_dp_wait_done), then the bookkeeping
    if (cae_last_block(a.ticket)) {
        if ((int)threadIdx.x < W) st_release_sys(a.peers.flags[threadIdx.x] + CAE_DP_MAX_WORLD + rank, e);
        if (threadIdx.x == 0) {
            *a.epoch = e;
            a.step_count[0] += 1;
            if (a.cursor) {
                const int c = a.cursor[0] + 1;
                a.cursor[0] = (c >= a.n_batches) ? 0 : c;
            }
        }
    }
}

// first launch of a step: every peer has finished reading this rank's gradients of the previous exchange
__global__ void k_dp_wait_done(const CaeDpPeers peers, const unsigned int* epoch) {
    if ((int)threadIdx.x < peers.world) {
        const unsigned int e = *reinterpret_cast<const volatile unsigned int*>(epoch);
        const unsigned int* f = peers.flags[peers.rank] + CAE_DP_MAX_WORLD + threadIdx.x;
        while ((int)(ld_acquire_sys(f) - e) < 0) {}
    }
}

extern "C" int cae_dp_wait_done(const CaeDpPeers* peers, const unsigned int* epoch, void* stream) {
    CAE_REQUIRE(peers && epoch && peers->world >= 2 && peers->world <= CAE_DP_MAX_WORLD, "dp_wait_done: bad argument");
    k_dp_wait_done<<<1, 32, 0, (cudaStream_t)stream>>>(*peers, epoch);
    return cae_check_launch("cae_dp_wait_done");
}

extern "C" int cae_adam_allreduce(float* p, const CaeDpPeers* peers, float* m, float* v, long long n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, int decoupled, float grad_scale, int* step_count,
                                  int* cursor, int n_batches, unsigned int* epoch, unsigned int* ticket, void* stream) {
    CAE_REQUIRE(p && peers && m && v && step_count && epoch && ticket && n > 0, "adam_allreduce: bad argument");
    CAE_REQUIRE(((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "adam_allreduce: arenas must be 16-byte aligned");
    CAE_REQUIRE(peers->world >= 2 && peers->world <= CAE_DP_MAX_WORLD && peers->rank >= 0 && peers->rank < peers->world,
                "adam_allreduce: world %d / rank %d outside 2..%d", peers->world, peers->rank, CAE_DP_MAX_WORLD);
    for (int r = 0; r < peers->world; ++r) CAE_REQUIRE(peers->grads[r] && peers->flags[r], "adam_allreduce: peer %d not mapped", r);
    DpArgs a;
    memset(&a, 0, sizeof(a));
    a.peers = *peers;
    a.p = p; a.m = m; a.v = v; a.n = n;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.gscale = grad_scale;
    a.decoupled = decoupled;
    a.step_count = step_count; a.cursor = cursor; a.n_batches = n_batches; a.epoch = epoch; a.ticket = ticket;
    // every CTA spins in phase 2: all of them must be resident at once (one wave)
    const int grid = (int)min((long long)CAE_NUM_SMS, (n + CAE_NT - 1) / CAE_NT);
    k_adam_allreduce<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_adam_allreduce");
}
