// Training-mode UNET stem (everything before the last transposed convolution) as ONE cooperative forward kernel and ONE
// cooperative backward kernel.  See include/cae_b200.h (CaeStemTrain) for the contract and the reference lines.
//
// Why: at batch 64 the stem moves < 1 MB and ~20 MFLOP per direction, yet as ~55 dependent launches it cost 257 of the
// 331 us of a training step - each launch paid cursor load -> staging -> compute -> partial row -> fence + ticket ->
// last-CTA finalize.  Here a CTA owns SPC (1..4) samples and keeps every activation (forward) / gradient (backward) of
// them in shared memory; the only cross-CTA traffic is what training-mode BatchNorm semantically requires: one
// per-channel (sum, sum of squares) row per CTA, a grid barrier, and a fixed-order column sum that every CTA repeats
// (bitwise identical everywhere, deterministic).  Weight gradients: one partial row per CTA, summed in row order.
//
// Inner loops: thread = (position, group of 4 output channels); the layer's weights are staged in shared memory as
// [ci][tap][co] so a tap costs one broadcast LDS.128 + one LDS per 4 FMAs.  Two generic routines cover the four
// convolution-like products: st_sconv (strided gather: Conv2d forward, ConvTranspose2d input gradient) and st_tconv
// (transposed gather: ConvTranspose2d forward, Conv2d input gradient); st_wgrad covers both weight gradients.
#include <cooperative_groups.h>
#include "capi_host.h"

namespace cg = cooperative_groups;

#ifndef ST_NT
#define ST_NT 256
#endif
#define ST_NW (ST_NT / 32)
#define ST_MAX_SPC 4
#define ST_CMAX 64            // channels per BatchNorm layer
#define ST_NBN 12             // coefficient tables: encoder l -> l, fc i -> 4 + i, decoder block j -> 8 + j

struct StPlan {
    // per-sample tape (floats): input, raw conv outputs, fc outputs, per block: raw convT output, raw concat, att, hid, pool
    int x, ye[CAE_STEM_MAX], t[CAE_STEM_MAX], yu[CAE_STEM_MAX], cat[CAE_STEM_MAX], att[CAE_STEM_MAX], hid[CAE_STEM_MAX],
        pool[CAE_STEM_MAX];
    int tape, act_max, w_max;
    int spc, ctas;
    // weight-gradient partial row
    int o_conv_w[CAE_STEM_MAX], o_fc_w[CAE_STEM_MAX], o_fc_b[CAE_STEM_MAX], o_up_w[CAE_STEM_MAX], o_up_b[CAE_STEM_MAX],
        o_up_w1[CAE_STEM_MAX], o_up_w2[CAE_STEM_MAX];
    int wrow;
    // backward: per-sample gradient arena (floats): dz (largest BatchNorm layer), two ping-pong buffers, skip gradients
    int g_dz, g_a, g_b, g_skip[CAE_STEM_MAX], garena;
    int pfloats;              // floats of the contiguous parameter block copied to shared memory
    int smem_fwd, smem_bwd;   // bytes
};

struct StArgs {
    CaeStemTrain s;
    CaeSrc x;
    StPlan p;
};

// ---- all stem parameters (weights, biases, BatchNorm gamma / beta: one contiguous block of the engine's flat arena) land in
// shared memory through ONE asynchronous bulk copy (cp.async.bulk + mbarrier) issued at kernel start: with 8 warps per CTA
// nothing hides an L2 round trip (~0.7 us), and the first version of this kernel spent most of its time in serialised
// parameter loads inside the layer loops (profiles/r02_stem_profile_v1.txt).
__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_params_fetch(float* dst, const float* src, int nfloats, uint64_t* bar) {
    if (threadIdx.x == 0) {
        const uint32_t b = st_smem_u32(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)nfloats * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        for (uint32_t o = 0; o < bytes; o += 32768u) {
            const uint32_t n = min(32768u, bytes - o);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(st_smem_u32(dst) + o), "l"((uint64_t)(reinterpret_cast<const char*>(src) + o)), "r"(n), "r"(b)
                         : "memory");
        }
    }
}
__device__ __forceinline__ void st_params_wait(uint64_t* bar) {
    const uint32_t b = st_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(b), "r"(0u)
            : "memory");
    } while (!done);
}

// phase timestamps of CTA 0 (clock64), read back through cae_unet_stem_train_profile: [0..31] forward, [32..63] backward
__device__ unsigned long long st_prof[64];
#define ST_T(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) st_prof[(i)] = clock64(); } while (0)

struct StDrop {
    uint32_t thresh, key;
    float scale;
};

// division by a run-time constant without the ~25-instruction integer divide: q = floor(x * ceil(2^32 / d) / 2^32), exact for
// x * d < 2^32 (every index here is < 2^20, every divisor < 2^12).  One 32-bit divide per construction.
struct StDiv {
    uint32_t m, d;
    // m: an over-estimate of 2^32 / d by at most 2^-21 relative (float reciprocal + guard), which keeps floor(x*m / 2^32)
    // exact for x < 2^20 - a handful of instructions instead of an integer divide per construction
    __device__ __forceinline__ explicit StDiv(int dd) : d((uint32_t)dd) {
        const uint32_t a = __float2uint_rz(4294967296.f * __frcp_rn((float)dd) * 0.99999994f);
        m = dd > 1 ? a + (a >> 21) + 2u : 0u;
    }
    __device__ __forceinline__ int div(int x) const { return d > 1 ? (int)__umulhi((uint32_t)x, m) : x; }
    __device__ __forceinline__ int mod(int x, int q) const { return x - q * (int)d; }
};

__device__ __forceinline__ uint32_t st_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// keep-mask of dropout site `site`, element `idx` (= sample * elements + element): counter-based, no state
__device__ __forceinline__ bool st_keep(const StDrop& d, uint32_t site, uint32_t idx) {
    if (d.thresh == 0u) return true;
    const uint32_t h = st_mix(idx + st_mix(d.key ^ (site * 0x85EBCA6Bu)));
    return (h >> 8) >= d.thresh;
}
__device__ __forceinline__ StDrop st_drop_init(const CaeStemTrain& s) {
    StDrop d;
    const float p = s.dropout_p;
    d.thresh = p > 0.f ? (uint32_t)(p * 16777216.f) : 0u;
    d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    const uint32_t step = s.step_count ? (uint32_t)__ldg(s.step_count) : 0u;
    d.key = st_mix((uint32_t)(s.seed & 0xffffffffu) ^ st_mix((uint32_t)(s.seed >> 32) + step * 0x9E3779B9u));
    return d;
}

// dst[e] = dropout(relu(scale[c] * src[e] + shift[c])), c = e / HW; scale == nullptr: no affine; site < 0: no dropout
__device__ __forceinline__ void st_make_act(float* dst, const float* src, int C, int HW, const float* scale, const float* shift,
                                            bool relu, const StDrop& d, int site, uint32_t base) {
    const int total = C * HW;
    const StDiv dHW(HW);
    for (int e = threadIdx.x; e < total; e += ST_NT) {
        float v = src[e];
        if (scale) {
            const int c = dHW.div(e);
            v = fmaf(v, scale[c], shift[c]);
        }
        if (relu) v = fmaxf(v, 0.f);
        if (site >= 0 && d.thresh) v = st_keep(d, (uint32_t)site, base + e) ? v * d.scale : 0.f;
        dst[e] = v;
    }
}

// weights stored [A][B][KK] -> shared [(ci*KK + t)*CoP + co]; ci_is_a: ci = a, co = b, else ci = b, co = a
__device__ __forceinline__ void st_stage_w(float* wsm, const float* w, int A, int B, int KK, bool ci_is_a) {
    const int Co = ci_is_a ? B : A, CoP = (Co + 3) & ~3, Ci = ci_is_a ? A : B;
    if (CoP != Co)                                           // zero the padding columns
        for (int e = threadIdx.x; e < Ci * KK * (CoP - Co); e += ST_NT) {
            const int r = e / (CoP - Co);                    // (rare: only when Cout is not a multiple of 4)
            wsm[r * CoP + Co + (e - r * (CoP - Co))] = 0.f;
        }
    const int total = A * B * KK;                            // `w` is the shared-memory copy of the parameter
    const StDiv dKK(KK), dB(B);
    for (int e = threadIdx.x; e < total; e += ST_NT) {
        const int r = dKK.div(e), t = dKK.mod(e, r), a_ = dB.div(r), b = dB.mod(r, a_);
        const int ci = ci_is_a ? a_ : b, co = ci_is_a ? b : a_;
        wsm[(ci * KK + t) * CoP + co] = w[e];
    }
}

// Thread layout of the two gather routines: item = (output position, group of 4 output channels); when a layer has fewer
// items than threads the input-channel reduction is split over KS adjacent lanes (power of two) and combined by a
// fixed-order butterfly - the deep narrow layers (32 x 2 x 2 ...) would otherwise run on a single warp.
__device__ __forceinline__ int st_ksplit(int items, int Ci) {
    int ks = 1;
    while (ks < 32 && items * ks * 2 <= ST_NT && ks * 2 <= Ci) ks *= 2;
    return ks;
}

// strided gather:  out[co][oy][ox] = bias[co] + sum_{ci,ky,kx} in[ci][oy*s - p + ky][ox*s - p + kx] * w[ci][t][co]
// K > 0: kernel size known at compile time (3 / 4: every spec of the reference) - the tap loops unroll into predicated
// LDS + LDS.128 + 4 FMA groups; K == 0: run-time loops.
template <int K>
__device__ __forceinline__ void st_sconv_k(const float* in, int Ci, int Hi, int Wi, float* out, int Co, int Ho, int Wo, int k, int s,
                                           int p, const float* wsm, const float* bias) {
    const int CoP = (Co + 3) & ~3, HWo = Ho * Wo, items = HWo * (CoP >> 2), KK = k * k;
    const int KS = st_ksplit(items, Ci), slice = threadIdx.x & (KS - 1), per = ST_NT / KS;
    const StDiv dHWo(HWo), dWo(Wo);
    const int HWi = Hi * Wi, wstep = KK * CoP;
    for (int it0 = 0; it0 < items; it0 += per) {
        const int it = it0 + threadIdx.x / KS;
        const bool valid = it < items;
        const int itc = valid ? it : 0;
        const int cgi = dHWo.div(itc), pos = dHWo.mod(itc, cgi), cg4 = cgi * 4, oy = dWo.div(pos), ox = dWo.mod(pos, oy);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int iy0 = oy * s - p, ix0 = ox * s - p;
        const int ky0 = max(0, -iy0), ky1 = valid ? min(k, Hi - iy0) : 0, kx0 = max(0, -ix0), kx1 = min(k, Wi - ix0);
        const float* ip = in + slice * HWi + iy0 * Wi + ix0;
        const float* wp = wsm + (size_t)slice * wstep + cg4;
        for (int ci = slice; ci < Ci; ci += KS, ip += KS * HWi, wp += KS * wstep) {
            if (K > 0) {
#pragma unroll
                for (int ky = 0; ky < K; ++ky) {
                    if (ky < ky0 || ky >= ky1) continue;
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        if (kx < kx0 || kx >= kx1) continue;
                        const float v = ip[ky * Wi + kx];
                        const float4 w4 = *reinterpret_cast<const float4*>(wp + (ky * K + kx) * CoP);
                        acc[0] = fmaf(v, w4.x, acc[0]); acc[1] = fmaf(v, w4.y, acc[1]);
                        acc[2] = fmaf(v, w4.z, acc[2]); acc[3] = fmaf(v, w4.w, acc[3]);
                    }
                }
            } else {
                for (int ky = ky0; ky < ky1; ++ky)
                    for (int kx = kx0; kx < kx1; ++kx) {
                        const float v = ip[ky * Wi + kx];
                        const float4 w4 = *reinterpret_cast<const float4*>(wp + (ky * k + kx) * CoP);
                        acc[0] = fmaf(v, w4.x, acc[0]); acc[1] = fmaf(v, w4.y, acc[1]);
                        acc[2] = fmaf(v, w4.z, acc[2]); acc[3] = fmaf(v, w4.w, acc[3]);
                    }
            }
        }
        for (int o = KS >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (valid && slice == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (cg4 + j < Co) out[(cg4 + j) * HWo + pos] = acc[j] + (bias ? bias[cg4 + j] : 0.f);
        }
    }
}
__device__ __forceinline__ void st_sconv(const float* in, int Ci, int Hi, int Wi, float* out, int Co, int Ho, int Wo, int k, int s,
                                         int p, const float* wsm, const float* bias) {
    if (k == 3) st_sconv_k<3>(in, Ci, Hi, Wi, out, Co, Ho, Wo, k, s, p, wsm, bias);
    else if (k == 4) st_sconv_k<4>(in, Ci, Hi, Wi, out, Co, Ho, Wo, k, s, p, wsm, bias);
    else st_sconv_k<0>(in, Ci, Hi, Wi, out, Co, Ho, Wo, k, s, p, wsm, bias);
}

// transposed gather:  out[co][oy][ox] = bias[co] + sum_{ci} sum_{ky = (oy+p) % s + m*s} in[ci][(oy+p-ky)/s][..] * w[ci][t][co]
// k <= 2*s (every transposed conv of the reference's specs: k4 s2, k3 s2): an output pixel has at most 2 x 2 taps; their
// input / weight offsets are computed once per item and the channel loop is four predicated LDS + LDS.128 + 4 FMA groups.
__device__ __forceinline__ void st_tconv(const float* in, int Ci, int Hi, int Wi, float* out, int Co, int Ho, int Wo, int k, int s,
                                         int p, const float* wsm, const float* bias) {
    const int CoP = (Co + 3) & ~3, HWo = Ho * Wo, items = HWo * (CoP >> 2), KK = k * k;
    const int KS = st_ksplit(items, Ci), slice = threadIdx.x & (KS - 1), per = ST_NT / KS;
    const StDiv dHWo(HWo), dWo(Wo), dS(s);
    const int HWi = Hi * Wi, wstep = KK * CoP;
    const bool two = k <= 2 * s;
    for (int it0 = 0; it0 < items; it0 += per) {
        const int it = it0 + threadIdx.x / KS;
        const bool valid = it < items;
        const int itc = valid ? it : 0;
        const int cgi = dHWo.div(itc), pos = dHWo.mod(itc, cgi), cg4 = cgi * 4, oy = dWo.div(pos), ox = dWo.mod(pos, oy);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // taps ky = kyf + m*s read input row iyf - m; valid m: ky < k, 0 <= iyf - m < Hi
        const int ty = oy + p, tx = ox + p;
        const int iyf = dS.div(ty), kyf = dS.mod(ty, iyf), ixf = dS.div(tx), kxf = dS.mod(tx, ixf);
        const float* ip = in + slice * HWi + iyf * Wi + ixf;
        const float* wp = wsm + ((size_t)slice * KK + kyf * k + kxf) * CoP + cg4;
        if (two) {
            bool ok[4];
            int io[4], wo[4];
#pragma unroll
            for (int my = 0; my < 2; ++my)
#pragma unroll
                for (int mx = 0; mx < 2; ++mx) {
                    const int u = my * 2 + mx;
                    ok[u] = valid && kyf + my * s < k && iyf - my >= 0 && iyf - my < Hi && kxf + mx * s < k && ixf - mx >= 0 &&
                            ixf - mx < Wi;
                    io[u] = -my * Wi - mx;
                    wo[u] = (my * s * k + mx * s) * CoP;
                }
            for (int ci = slice; ci < Ci; ci += KS, ip += KS * HWi, wp += KS * wstep) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok[u]) {
                        const float v = ip[io[u]];
                        const float4 w4 = *reinterpret_cast<const float4*>(wp + wo[u]);
                        acc[0] = fmaf(v, w4.x, acc[0]); acc[1] = fmaf(v, w4.y, acc[1]);
                        acc[2] = fmaf(v, w4.z, acc[2]); acc[3] = fmaf(v, w4.w, acc[3]);
                    }
            }
        } else {
            const int my0 = max(0, iyf - Hi + 1), my1 = valid ? min(dS.div(k - kyf + s - 1), iyf + 1) : 0;
            const int mx0 = max(0, ixf - Wi + 1), mx1 = min(dS.div(k - kxf + s - 1), ixf + 1);
            for (int ci = slice; ci < Ci; ci += KS, ip += KS * HWi, wp += KS * wstep)
                for (int my = my0; my < my1; ++my)
                    for (int mx = mx0; mx < mx1; ++mx) {
                        const float v = ip[-my * Wi - mx];
                        const float4 w4 = *reinterpret_cast<const float4*>(wp + (my * s * k + mx * s) * CoP);
                        acc[0] = fmaf(v, w4.x, acc[0]); acc[1] = fmaf(v, w4.y, acc[1]);
                        acc[2] = fmaf(v, w4.z, acc[2]); acc[3] = fmaf(v, w4.w, acc[3]);
                    }
        }
        for (int o = KS >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (valid && slice == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (cg4 + j < Co) out[(cg4 + j) * HWo + pos] = acc[j] + (bias ? bias[cg4 + j] : 0.f);
        }
    }
}

// weight gradient of both convolutions: dW[a][b][ky][kx] (+)= sum_{sy,sx} small[a][sy][sx] * big[b][sy*s - p + ky][sx*s - p + kx]
//   ConvTranspose2d W[ci][co]: small = layer input (a = ci), big = dL/dy (b = co)
//   Conv2d          W[co][ci]: small = dL/dy (a = co),       big = layer input (b = ci)
// K = 3 / 4 (every layer of the reference's specs): one (a, b) channel pair per thread with all K*K taps in registers - each
// small-tensor value is loaded once for K*K FMAs; layers with few pairs split the rows of the small tensor over PS lanes.
template <int K>
__device__ __forceinline__ void st_wgrad_k(const float* sm_, int Hs, int Ws, const float* bg, int Hb, int Wb, int A, int B, int s,
                                           int p, float* dst, bool accumulate) {
    const int pairs = A * B;
    int PS = 1;
    while (PS < 32 && pairs * PS * 2 <= ST_NT && PS * 2 <= Hs) PS *= 2;
    const int slice = threadIdx.x & (PS - 1), per = ST_NT / PS;
    const StDiv dB(B);
    for (int q0 = 0; q0 < pairs; q0 += per) {
        const int q = q0 + threadIdx.x / PS;
        const bool valid = q < pairs;
        const int qc = valid ? q : 0;
        const int a = dB.div(qc), b = dB.mod(qc, a);
        const float* sp = sm_ + a * Hs * Ws;
        const float* bp = bg + b * Hb * Wb;
        float acc[K * K];
#pragma unroll
        for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
        for (int sy = valid ? slice : Hs; sy < Hs; sy += PS) {
            const int by0 = sy * s - p;
            for (int sx = 0; sx < Ws; ++sx) {
                const float v = sp[sy * Ws + sx];
                const int bx0 = sx * s - p;
#pragma unroll
                for (int ky = 0; ky < K; ++ky) {
                    const int by = by0 + ky;
                    if (by < 0 || by >= Hb) continue;
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        const int bx = bx0 + kx;
                        if (bx >= 0 && bx < Wb) acc[ky * K + kx] = fmaf(v, bp[by * Wb + bx], acc[ky * K + kx]);
                    }
                }
            }
        }
        for (int o = PS >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int t = 0; t < K * K; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        }
        if (valid && slice == 0) {
            float* d = dst + (size_t)q * (K * K);
#pragma unroll
            for (int t = 0; t < K * K; ++t) d[t] = accumulate ? d[t] + acc[t] : acc[t];
        }
    }
}

__device__ __forceinline__ void st_wgrad_generic(const float* sm_, int Hs, int Ws, const float* bg, int Hb, int Wb, int A, int B, int k,
                                                 int s, int p, float* dst, bool accumulate) {
    const int KK = k * k, total = A * B * KK;
    int PS = 1;
    while (PS < 32 && total * PS * 2 <= ST_NT && PS * 2 <= Hs) PS *= 2;
    const int slice = threadIdx.x & (PS - 1), per = ST_NT / PS;
    for (int e0 = 0; e0 < total; e0 += per) {
        const int e = e0 + threadIdx.x / PS;
        const bool valid = e < total;
        const int ec = valid ? e : 0;
        const int t = ec % KK, r = ec / KK, b = r % B, a = r / B, ky = t / k, kx = t - ky * k;
        const float* sp = sm_ + a * Hs * Ws;
        const float* bp = bg + b * Hb * Wb + (ky - p) * Wb + (kx - p);
        // rows / columns of the small tensor whose partner lies inside the big one: 0 <= sy*s - p + ky < Hb
        const int sy0 = max(0, (p - ky + s - 1) / s), sy1 = (valid && Hb - 1 + p - ky >= 0) ? min(Hs, (Hb - 1 + p - ky) / s + 1) : 0;
        const int sx0 = max(0, (p - kx + s - 1) / s), sx1 = (Wb - 1 + p - kx >= 0) ? min(Ws, (Wb - 1 + p - kx) / s + 1) : 0;
        float acc = 0.f;
        for (int sy = sy0 + slice; sy < sy1; sy += PS)
            for (int sx = sx0; sx < sx1; ++sx) acc = fmaf(sp[sy * Ws + sx], bp[sy * s * Wb + sx * s], acc);
        for (int o = PS >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (valid && slice == 0) dst[e] = accumulate ? dst[e] + acc : acc;
    }
}

__device__ __forceinline__ void st_wgrad(const float* sm_, int Hs, int Ws, const float* bg, int Hb, int Wb, int A, int B, int k, int s,
                                         int p, float* dst, bool accumulate) {
    if (k == 3) st_wgrad_k<3>(sm_, Hs, Ws, bg, Hb, Wb, A, B, s, p, dst, accumulate);
    else if (k == 4) st_wgrad_k<4>(sm_, Hs, Ws, bg, Hb, Wb, A, B, s, p, dst, accumulate);
    else st_wgrad_generic(sm_, Hs, Ws, bg, Hb, Wb, A, B, k, s, p, dst, accumulate);
}

// per-channel (sum, sum of squares) of one sample's [C][HW] tensor, added to the CTA accumulators (double)
__device__ __forceinline__ void st_chan_stats(const float* v, int C, int HW, double* sacc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (HW == 1) {
        for (int c = threadIdx.x; c < C; c += ST_NT) {
            const double x = (double)v[c];
            sacc[2 * c] += x;
            sacc[2 * c + 1] += x * x;
        }
        return;
    }
    for (int c = warp; c < C; c += ST_NW) {
        float s = 0.f, q = 0.f;
        for (int i = lane; i < HW; i += 32) {
            const float x = v[c * HW + i];
            s += x;
            q = fmaf(x, x, q);
        }
        const double S = warp_sum_d((double)s), Q = warp_sum_d((double)q);
        if (lane == 0) {
            sacc[2 * c] += S;
            sacc[2 * c + 1] += Q;
        }
    }
}

// backward statistics of one sample: S1 += sum dz, S2 += sum dz * xhat, xhat = (raw - mean) * invstd
__device__ __forceinline__ void st_chan_stats_bwd(const float* dz, const float* raw, int C, int HW, const float* mean,
                                                  const float* invstd, double* sacc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (HW == 1) {
        for (int c = threadIdx.x; c < C; c += ST_NT) {
            const double d = (double)dz[c];
            sacc[2 * c] += d;
            sacc[2 * c + 1] += d * (double)((raw[c] - mean[c]) * invstd[c]);
        }
        return;
    }
    for (int c = warp; c < C; c += ST_NW) {
        float s = 0.f, q = 0.f;
        const float m = mean[c], is = invstd[c];
        for (int i = lane; i < HW; i += 32) {
            const float d = dz[c * HW + i];
            s += d;
            q = fmaf(d, (raw[c * HW + i] - m) * is, q);
        }
        const double S = warp_sum_d((double)s), Q = warp_sum_d((double)q);
        if (lane == 0) {
            sacc[2 * c] += S;
            sacc[2 * c + 1] += Q;
        }
    }
}

// Cross-CTA column sums of the per-CTA rows: every CTA publishes its 2*C doubles, the grid meets, every CTA adds the rows
// in row order (the same order everywhere: identical results in all CTAs).  tot[2*c], tot[2*c+1] in shared memory.
// `parity` alternates the scratch half so a fast CTA's next publication cannot overwrite rows a slow CTA still reads.
__device__ __forceinline__ void st_grid_sums(cg::grid_group& grid, double* bnpart, int parity, int C, const double* sacc, double* tot,
                                             int& ti) {
    ST_T(ti++);                                              // compute of this phase done
    double* mine = bnpart + ((size_t)parity * gridDim.x + blockIdx.x) * (2 * ST_CMAX);
    for (int i = threadIdx.x; i < 2 * C; i += ST_NT) mine[i] = sacc[i];
    __threadfence();
    grid.sync();
    ST_T(ti++);                                              // barrier passed
    const double* base = bnpart + (size_t)parity * gridDim.x * (2 * ST_CMAX);
    const int rows = gridDim.x;
    // 2*C columns; TPC threads per column (power of two, groups inside a warp), fixed-order butterfly
    int tpc = 1;
    while (tpc < 32 && tpc * 2 * (2 * C) <= ST_NT) tpc *= 2;
    const int col = threadIdx.x / tpc, sub = threadIdx.x & (tpc - 1);
    double s = 0.0;
    if (col < 2 * C) {
        // loads issued sixteen (then four) at a time - one exposed L2 latency per batch - and added in row order
        const double* cp = base + col;
        const size_t rs = 2 * ST_CMAX;
        int r = sub;
        for (; r + 15 * tpc < rows; r += 16 * tpc) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldcg(cp + (size_t)(r + u * tpc) * rs);
#pragma unroll
            for (int u = 0; u < 16; ++u) s += v[u];
        }
        for (; r + 3 * tpc < rows; r += 4 * tpc) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldcg(cp + (size_t)(r + u * tpc) * rs);
#pragma unroll
            for (int u = 0; u < 4; ++u) s += v[u];
        }
        for (; r < rows; r += tpc) s += __ldcg(cp + (size_t)r * rs);
    }
    for (int o = tpc >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (col < 2 * C && sub == 0) tot[col] = s;
    __syncthreads();
    ST_T(ti++);                                              // column sums done
}

// forward BatchNorm finalisation from the grid totals: coefficient tables (shared) + module buffers (CTA 0, global)
__device__ __forceinline__ void st_bn_finalize_fwd(const CaeBN& bn, const float* gamma, const float* beta, const double* tot,
                                                   double count, float* coef) {
    for (int c = threadIdx.x; c < bn.C; c += ST_NT) {
        const double mean = tot[2 * c] / count;
        double var = tot[2 * c + 1] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const double invstd = rsqrt(var + (double)bn.eps);
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        const float scale = (float)((double)g * invstd), shift = (float)((double)b - mean * (double)g * invstd);
        coef[c] = scale;
        coef[ST_CMAX + c] = shift;
        coef[2 * ST_CMAX + c] = (float)mean;
        coef[3 * ST_CMAX + c] = (float)invstd;
        if (blockIdx.x == 0) {
            bn.scale[c] = scale;
            bn.shift[c] = shift;
            bn.mean[c] = (float)mean;
            bn.invstd[c] = (float)invstd;
            if (bn.running_mean) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                const double m = (double)bn.momentum;
                bn.running_mean[c] = (float)((1.0 - m) * (double)bn.running_mean[c] + m * mean);
                bn.running_var[c] = (float)((1.0 - m) * (double)bn.running_var[c] + m * unbiased);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && bn.num_batches_tracked) bn.num_batches_tracked[0] += 1;
    __syncthreads();
}

// backward: dL/dy = A * dz + B * raw + C per channel; bc = [A | B | C] (shared); CTA 0 writes dgamma / dbeta / dead bias
__device__ __forceinline__ void st_bn_finalize_bwd(const CaeBN& bn, const float* gamma, const double* tot, double count,
                                                   const float* coef, float* bc) {
    for (int c = threadIdx.x; c < bn.C; c += ST_NT) {
        const double S1 = tot[2 * c], S2 = tot[2 * c + 1];
        const double g = gamma ? (double)gamma[c] : 1.0;
        const double invstd = (double)coef[3 * ST_CMAX + c], mean = (double)coef[2 * ST_CMAX + c];
        const double A = g * invstd, B = -A * invstd * S2 / count, Cc = -A * S1 / count - B * mean;
        bc[c] = (float)A;
        bc[ST_CMAX + c] = (float)B;
        bc[2 * ST_CMAX + c] = (float)Cc;
        if (blockIdx.x == 0) {
            if (bn.dgamma) bn.dgamma[c] = (float)S2;
            if (bn.dbeta) bn.dbeta[c] = (float)S1;
            if (bn.dbias) bn.dbias[c] = 0.f;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void st_zero_acc(double* sacc, int n) {
    for (int i = threadIdx.x; i < n; i += ST_NT) sacc[i] = 0.0;
}

// one fully connected layer of one sample: out[o] = b[o] + sum_k W[o][k] in[k]  (W, b: shared-memory copies)
__device__ __forceinline__ void st_fc(const float* W, const float* b, const float* in, int K, float* out, int O, bool relu) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (K >= 32) {
        for (int o = warp; o < O; o += ST_NW) {
            float acc = 0.f;
            for (int k = lane; k < K; k += 32) acc = fmaf(W[(size_t)o * K + k], in[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                acc += b ? b[o] : 0.f;
                out[o] = relu ? fmaxf(acc, 0.f) : acc;
            }
        }
    } else {
        for (int o = threadIdx.x; o < O; o += ST_NT) {
            float acc = b ? b[o] : 0.f;
            for (int k = 0; k < K; ++k) acc = fmaf(W[(size_t)o * K + k], in[k], acc);
            out[o] = relu ? fmaxf(acc, 0.f) : acc;
        }
    }
}

// din[k] = sum_o dpre[o] W[o][k]
__device__ __forceinline__ void st_fc_dx(const float* W, const float* dpre, int O, float* din, int K) {
    for (int k = threadIdx.x; k < K; k += ST_NT) {
        float acc = 0.f;
        for (int o = 0; o < O; ++o) acc = fmaf(dpre[o], W[(size_t)o * K + k], acc);
        din[k] = acc;
    }
}

// dW[o][k] (+)= dpre[o] * in[k] ; db[o] (+)= dpre[o]
__device__ __forceinline__ void st_fc_dw(const float* dpre, int O, const float* in, int K, float* dW, float* db, bool accumulate) {
    for (int e = threadIdx.x; e < O * K; e += ST_NT) {
        const int o = e / K, k = e - o * K;
        const float v = dpre[o] * in[k];
        dW[e] = accumulate ? dW[e] + v : v;
    }
    if (db)
        for (int o = threadIdx.x; o < O; o += ST_NT) db[o] = accumulate ? db[o] + dpre[o] : dpre[o];
}

#define ST_COEF(bnidx) (coef + (bnidx) * 4 * ST_CMAX)

// ============================================================================================================
// forward
// ============================================================================================================
__global__ void __launch_bounds__(ST_NT, 1) k_unet_stem_train_fwd(const StArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::grid_group grid = cg::this_grid();
    const CaeStemTrain& S = a.s;
    const StPlan& P = a.p;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * P.spc;
    const int ns = max(0, min(P.spc, S.N - n0));
    float* tape = sm;                                       // [spc][P.tape]
    float* act = tape + P.spc * P.tape;                     // [act_max]
    float* wsm = act + P.act_max;                           // [w_max]
    float* coef = wsm + P.w_max;                            // [ST_NBN][4][ST_CMAX]
    double* sacc = reinterpret_cast<double*>(coef + ST_NBN * 4 * ST_CMAX);   // [2*ST_CMAX]
    double* tot = sacc + 2 * ST_CMAX;                       // [2*ST_CMAX]
    uint64_t* pbar = reinterpret_cast<uint64_t*>(tot + 2 * ST_CMAX);
    float* prm = reinterpret_cast<float*>(pbar + 2);        // [P.pfloats] shared copy of the stem parameters
#define PRM(ptr) ((ptr) ? prm + ((ptr) - S.params) : nullptr)
    st_params_fetch(prm, S.params, P.pfloats, pbar);
    const StDrop drop = st_drop_init(S);
    int parity = 0, ti = 0;
    ST_T(ti++);
#define TP(si, off) (tape + (si) * P.tape + (off))

    // ---- input
    {
        const CaeStemTrainConv& c0 = S.conv[0];
        const int in_elems = c0.Cin * c0.Hin * c0.Win;
        const CaeView& xv = a.x.t0;
        const long long xbase = src_cursor_offset(a.x);
        for (int e = tid; e < ns * in_elems; e += ST_NT) {
            const int si = e / in_elems, r = e - si * in_elems;
            const int c = r / (c0.Hin * c0.Win), q = r - c * c0.Hin * c0.Win, yy = q / c0.Win, xx = q - yy * c0.Win;
            const ChanCoef kc = load_coef(a.x, c);
            TP(si, P.x)[r] = src_value(a.x, xbase + (long long)(n0 + si) * xv.sN + (long long)c * xv.sC + (long long)yy * xv.ld + xx, kc);
        }
    }
    __syncthreads();               // (the barrier was initialised by thread 0 above)
    st_params_wait(pbar);
    // ---- encoder: Conv2d -> BatchNorm2d (batch statistics) -> ReLU -> Dropout
    for (int l = 0; l < S.n_conv; ++l) {
        const CaeStemTrainConv& L = S.conv[l];
        const int HWo = L.Hout * L.Wout, in_elems = L.Cin * L.Hin * L.Win;
        st_stage_w(wsm, PRM(L.w), L.Cout, L.Cin, L.k * L.k, false);
        st_zero_acc(sacc, 2 * L.Cout);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const float* in = TP(si, P.x);
            if (l > 0) {
                const CaeStemTrainConv& Lp = S.conv[l - 1];
                st_make_act(act, TP(si, P.ye[l - 1]), Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(l - 1), ST_COEF(l - 1) + ST_CMAX, true,
                            drop, l - 1, (uint32_t)(n0 + si) * (uint32_t)in_elems);
                in = act;
                __syncthreads();
            }
            st_sconv(in, L.Cin, L.Hin, L.Win, TP(si, P.ye[l]), L.Cout, L.Hout, L.Wout, L.k, L.stride, L.pad, wsm, PRM(L.b));
            __syncthreads();
            st_chan_stats(TP(si, P.ye[l]), L.Cout, HWo, sacc);
            __syncthreads();
        }
        st_grid_sums(grid, S.bnpart, parity, L.Cout, sacc, tot, ti);
        parity ^= 1;
        st_bn_finalize_fwd(L.bn, PRM(L.bn.gamma), PRM(L.bn.beta), tot, (double)S.N * HWo, ST_COEF(l));
    }
    // ---- fc stacks: Linear [-> BatchNorm1d] -> ReLU -> Dropout
    for (int i = 0; i < S.n_fc; ++i) {
        const CaeStemTrainFc& L = S.fc[i];
        if (L.has_bn) st_zero_acc(sacc, 2 * L.out);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const uint32_t base = (uint32_t)(n0 + si) * (uint32_t)L.in;
            if (i == 0) {
                const CaeStemTrainConv& Lp = S.conv[S.n_conv - 1];
                st_make_act(act, TP(si, P.ye[S.n_conv - 1]), Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(S.n_conv - 1),
                            ST_COEF(S.n_conv - 1) + ST_CMAX, true, drop, S.n_conv - 1, base);
            } else if (S.fc[i - 1].has_bn) {
                st_make_act(act, TP(si, P.t[i - 1]), L.in, 1, ST_COEF(4 + i - 1), ST_COEF(4 + i - 1) + ST_CMAX, true, drop, 4 + i - 1, base);
            } else {
                st_make_act(act, TP(si, P.t[i - 1]), L.in, 1, nullptr, nullptr, false, drop, 4 + i - 1, base);   // stored post-ReLU
            }
            __syncthreads();
            st_fc(PRM(L.w), PRM(L.b), act, L.in, TP(si, P.t[i]), L.out, !L.has_bn);
            __syncthreads();
            if (L.has_bn) {
                st_chan_stats(TP(si, P.t[i]), L.out, 1, sacc);
                __syncthreads();
            }
        }
        if (L.has_bn) {
            st_grid_sums(grid, S.bnpart, parity, L.out, sacc, tot, ti);
            parity ^= 1;
            st_bn_finalize_fwd(L.bn, PRM(L.bn.gamma), PRM(L.bn.beta), tot, (double)S.N, ST_COEF(4 + i));
        }
    }
    // ---- decoder blocks: ConvTranspose2d -> ChannelAttention gate -> concat(skip) -> BatchNorm2d(2C) -> ReLU -> Dropout
    for (int j = 0; j < S.n_up; ++j) {
        const CaeStemTrainUp& L = S.up[j];
        const int C = L.Cout, HW = L.Hout * L.Wout, Cr = L.Cr, in_elems = L.Cin * L.Hin * L.Win;
        st_stage_w(wsm, PRM(L.w), L.Cin, L.Cout, L.k * L.k, true);
        const float* W1 = PRM(L.W1);
        const float* W2 = PRM(L.W2);
        st_zero_acc(sacc, 4 * C);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const uint32_t base = (uint32_t)(n0 + si) * (uint32_t)in_elems;
            if (j == 0) {
                const CaeStemTrainFc& F = S.fc[S.n_fc - 1];
                if (F.has_bn) st_make_act(act, TP(si, P.t[S.n_fc - 1]), in_elems, 1, ST_COEF(4 + S.n_fc - 1), ST_COEF(4 + S.n_fc - 1) + ST_CMAX,
                                          true, drop, 4 + S.n_fc - 1, base);
                else st_make_act(act, TP(si, P.t[S.n_fc - 1]), in_elems, 1, nullptr, nullptr, false, drop, 4 + S.n_fc - 1, base);
            } else {
                const CaeStemTrainUp& Lp = S.up[j - 1];
                st_make_act(act, TP(si, P.cat[j - 1]), 2 * Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(8 + j - 1), ST_COEF(8 + j - 1) + ST_CMAX, true,
                            drop, 8 + j - 1, base);
            }
            __syncthreads();
            float* y = TP(si, P.yu[j]);
            st_tconv(act, L.Cin, L.Hin, L.Win, y, C, L.Hout, L.Wout, L.k, L.stride, L.pad, wsm, PRM(L.b));
            __syncthreads();
            // plane statistics (avg, max, first arg-max): one warp per channel
            float* pool = TP(si, P.pool[j]);          // [avg C | max C | argmax C]
            {
                const int lane = tid & 31, warp = tid >> 5;
                for (int c = warp; c < C; c += ST_NW) {
                    float s = 0.f, m = -INFINITY;
                    int am = 0x7fffffff;
                    for (int i2 = lane; i2 < HW; i2 += 32) {
                        const float v = y[c * HW + i2];
                        s += v;
                        if (v > m) { m = v; am = i2; }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
                        const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
                        if (m2 > m || (m2 == m && a2 < am)) { m = m2; am = a2; }
                    }
                    const double Sd = warp_sum_d((double)s);
                    if (lane == 0) {
                        pool[c] = (float)(Sd / HW);
                        pool[C + c] = m;
                        pool[2 * C + c] = (float)am;
                    }
                }
            }
            __syncthreads();
            float* hid = TP(si, P.hid[j]);            // [2][Cr]
            for (int i2 = tid; i2 < 2 * Cr; i2 += ST_NT) {
                const int which = i2 / Cr, r = i2 - which * Cr;
                const float* src = pool + which * C;
                float v = 0.f;
                for (int c = 0; c < C; ++c) v = fmaf(W1[r * C + c], src[c], v);
                hid[i2] = fmaxf(v, 0.f);
            }
            __syncthreads();
            float* att = TP(si, P.att[j]);
            for (int c = tid; c < C; c += ST_NT) {
                float v = 0.f;
                for (int r = 0; r < Cr; ++r) v = fmaf(W2[c * Cr + r], hid[r] + hid[Cr + r], v);
                att[c] = 1.f / (1.f + expf(-v));
            }
            __syncthreads();
            // cat = [att * y ; relu(bn(skip))]  (the skip is the activation BEFORE dropout)
            float* cat = TP(si, P.cat[j]);
            const float* sk = TP(si, P.ye[L.skip]);
            const float* ssc = ST_COEF(L.skip);
            const StDiv dHW(HW);
            for (int e = tid; e < 2 * C * HW; e += ST_NT) {
                const int c2 = dHW.div(e);
                float v;
                if (c2 < C) v = att[c2] * y[e];
                else v = fmaxf(fmaf(sk[e - C * HW], ssc[c2 - C], ssc[ST_CMAX + c2 - C]), 0.f);
                cat[e] = v;
            }
            __syncthreads();
            st_chan_stats(cat, 2 * C, HW, sacc);
            __syncthreads();
        }
        st_grid_sums(grid, S.bnpart, parity, 2 * C, sacc, tot, ti);
        parity ^= 1;
        st_bn_finalize_fwd(L.bn, PRM(L.bn.gamma), PRM(L.bn.beta), tot, (double)S.N * HW, ST_COEF(8 + j));
    }
    // ---- activated input of the head + the tape
    {
        const CaeStemTrainUp& L = S.up[S.n_up - 1];
        const int elems = 2 * L.Cout * L.Hout * L.Wout;
        for (int si = 0; si < ns; ++si) {
            st_make_act(S.hin + (size_t)(n0 + si) * elems, TP(si, P.cat[S.n_up - 1]), 2 * L.Cout, L.Hout * L.Wout, ST_COEF(8 + S.n_up - 1),
                        ST_COEF(8 + S.n_up - 1) + ST_CMAX, true, drop, 8 + S.n_up - 1, (uint32_t)(n0 + si) * (uint32_t)elems);
        }
        for (int e = tid; e < ns * P.tape; e += ST_NT) S.tape[(size_t)n0 * P.tape + e] = tape[e];
    }
    ST_T(ti++);
    if (blockIdx.x == 0 && threadIdx.x == 0) st_prof[31] = (unsigned long long)ti;
#undef TP
}

// ============================================================================================================
// backward
// ============================================================================================================
__global__ void __launch_bounds__(ST_NT, 1) k_unet_stem_train_bwd(const StArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::grid_group grid = cg::this_grid();
    const CaeStemTrain& S = a.s;
    const StPlan& P = a.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * P.spc;
    const int ns = max(0, min(P.spc, S.N - n0));
    float* tape = sm;                                       // [spc][P.tape]
    float* gar = tape + P.spc * P.tape;                     // [spc][P.garena]
    float* act = gar + P.spc * P.garena;                    // [act_max]
    float* wsm = act + P.act_max;                           // [w_max]
    float* coef = wsm + P.w_max;                            // [ST_NBN][4][ST_CMAX]
    float* bc = coef + ST_NBN * 4 * ST_CMAX;                // [3][ST_CMAX]
    float* small = bc + 3 * ST_CMAX;                        // attention scratch [6*ST_CMAX]
    double* sacc = reinterpret_cast<double*>(small + 6 * ST_CMAX);
    double* tot = sacc + 2 * ST_CMAX;
    uint64_t* pbar = reinterpret_cast<uint64_t*>(tot + 2 * ST_CMAX);
    float* prm = reinterpret_cast<float*>(pbar + 2);        // [P.pfloats]
    st_params_fetch(prm, S.params, P.pfloats, pbar);
    const StDrop drop = st_drop_init(S);
    float* wrow = S.wpart + (size_t)blockIdx.x * P.wrow;
    int parity = 0, ti = 32;
    ST_T(ti++);
#define TP(si, off) (tape + (si) * P.tape + (off))
#define GA(si, off) (gar + (si) * P.garena + (off))

    // ---- tape + BatchNorm coefficients (written by the forward kernel)
    for (int e = tid; e < ns * P.tape; e += ST_NT) tape[e] = S.tape[(size_t)n0 * P.tape + e];
    for (int e = tid; e < ST_NBN * ST_CMAX; e += ST_NT) {
        const int b = e / ST_CMAX, c = e - b * ST_CMAX;
        // (fields copied by value: taking the address of a kernel-parameter member would force a local copy of it)
        const float *psc = nullptr, *psh = nullptr, *pme = nullptr, *pis = nullptr;
        int bC = 0;
        if (b < 4) {
            if (b < S.n_conv) { bC = S.conv[b].bn.C; psc = S.conv[b].bn.scale; psh = S.conv[b].bn.shift; pme = S.conv[b].bn.mean; pis = S.conv[b].bn.invstd; }
        } else if (b < 8) {
            if (b - 4 < S.n_fc && S.fc[b - 4].has_bn) { bC = S.fc[b - 4].bn.C; psc = S.fc[b - 4].bn.scale; psh = S.fc[b - 4].bn.shift; pme = S.fc[b - 4].bn.mean; pis = S.fc[b - 4].bn.invstd; }
        } else if (b - 8 < S.n_up) {
            bC = S.up[b - 8].bn.C; psc = S.up[b - 8].bn.scale; psh = S.up[b - 8].bn.shift; pme = S.up[b - 8].bn.mean; pis = S.up[b - 8].bn.invstd;
        }
        if (c < bC) {
            coef[b * 4 * ST_CMAX + c] = psc[c];
            coef[b * 4 * ST_CMAX + ST_CMAX + c] = psh[c];
            coef[b * 4 * ST_CMAX + 2 * ST_CMAX + c] = pme[c];
            coef[b * 4 * ST_CMAX + 3 * ST_CMAX + c] = pis[c];
        }
    }
    // CTAs without samples still publish zero rows
    for (int e = tid; ns == 0 && e < P.wrow; e += ST_NT) wrow[e] = 0.f;
    __syncthreads();
    st_params_wait(pbar);
    ST_T(ti++);

    // dz = g * dropout-mask * [bn(raw) > 0]  for a BatchNorm + ReLU + Dropout output; g may alias dz
    auto mask_to_dz = [&](float* dz, const float* g, const float* raw, int C, int HW, const float* cf, int site, uint32_t base,
                          const float* addend) {
        const StDiv dHW(HW);
        for (int e = tid; e < C * HW; e += ST_NT) {
            const int c = dHW.div(e);
            const float z = fmaf(raw[e], cf[c], cf[ST_CMAX + c]);
            float v = g[e];
            if (drop.thresh) v = st_keep(drop, (uint32_t)site, base + e) ? v * drop.scale : 0.f;
            if (addend) v += addend[e];              // gradient arriving through the skip connection (before dropout)
            dz[e] = z > 0.f ? v : 0.f;
        }
    };
    // in place: dy = A * dz + B * raw + C
    auto bn_apply_bwd = [&](float* dz, const float* raw, int C, int HW) {
        const StDiv dHW(HW);
        for (int e = tid; e < C * HW; e += ST_NT) {
            const int c = dHW.div(e);
            dz[e] = fmaf(dz[e], bc[c], fmaf(raw[e], bc[ST_CMAX + c], bc[2 * ST_CMAX + c]));
        }
    };

    // ---- head gradient -> dz of the last block's BatchNorm
    {
        const int j = S.n_up - 1;
        const CaeStemTrainUp& L = S.up[j];
        const int C2 = 2 * L.Cout, HW = L.Hout * L.Wout, elems = C2 * HW;
        st_zero_acc(sacc, 2 * C2);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            float* dz = GA(si, P.g_dz);
            for (int e = tid; e < elems; e += ST_NT) dz[e] = __ldg(S.dhin + (size_t)(n0 + si) * elems + e);
            __syncthreads();
            mask_to_dz(dz, dz, TP(si, P.cat[j]), C2, HW, ST_COEF(8 + j), 8 + j, (uint32_t)(n0 + si) * (uint32_t)elems, nullptr);
            __syncthreads();
            st_chan_stats_bwd(dz, TP(si, P.cat[j]), C2, HW, ST_COEF(8 + j) + 2 * ST_CMAX, ST_COEF(8 + j) + 3 * ST_CMAX, sacc);
            __syncthreads();
        }
        st_grid_sums(grid, S.bnpart, parity, C2, sacc, tot, ti);
        parity ^= 1;
        st_bn_finalize_bwd(L.bn, PRM(L.bn.gamma), tot, (double)S.N * HW, ST_COEF(8 + j), bc);
    }
    // ---- decoder blocks, last to first; block 0 continues into the fc stacks up to the next BatchNorm
    for (int j = S.n_up - 1; j >= 0; --j) {
        const CaeStemTrainUp& L = S.up[j];
        const int C = L.Cout, HW = L.Hout * L.Wout, Cr = L.Cr, in_elems = L.Cin * L.Hin * L.Win, KK = L.k * L.k;
        // input gradient of the transposed conv = strided gather over dy with w[co][t][ci]
        st_stage_w(wsm, PRM(L.w), L.Cin, L.Cout, KK, false);
        const float* W1 = PRM(L.W1);
        const float* W2 = PRM(L.W2);
        // what the next BatchNorm reduction is: block j-1's BN(2C), or (j == 0) the BatchNorm1d of fc[n_fc-2]
        const int next_C = j > 0 ? 2 * S.up[j - 1].Cout : S.fc[S.n_fc - 2].out;
        st_zero_acc(sacc, 2 * next_C);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const bool accw = si > 0;
            float* dcat = GA(si, P.g_dz);
            const float* cat = TP(si, P.cat[j]);
            const float* y = TP(si, P.yu[j]);
            const float* att = TP(si, P.att[j]);
            const float* hid = TP(si, P.hid[j]);
            const float* pool = TP(si, P.pool[j]);
            bn_apply_bwd(dcat, cat, 2 * C, HW);
            __syncthreads();
            // skip half -> gradient of the encoder activation (before its dropout)
            float* dsk = GA(si, P.g_skip[L.skip]);
            for (int e = tid; e < C * HW; e += ST_NT) dsk[e] = dcat[C * HW + e];
            // gate backward (ChannelAttention): ds = (sum g*y) * att * (1 - att)
            float* ds = small;                 // [C]
            float* dh = small + ST_CMAX;       // [2][Cr]
            float* dav = small + 2 * ST_CMAX;  // [C]
            float* dmx = small + 3 * ST_CMAX;  // [C]
            for (int c = warp; c < C; c += ST_NW) {
                float s = 0.f;
                for (int i2 = lane; i2 < HW; i2 += 32) s = fmaf(dcat[c * HW + i2], y[c * HW + i2], s);
                const double Sd = warp_sum_d((double)s);
                if (lane == 0) ds[c] = (float)Sd * att[c] * (1.f - att[c]);
            }
            __syncthreads();
            for (int i2 = tid; i2 < 2 * Cr; i2 += ST_NT) {
                const int r = i2 % Cr;
                float v = 0.f;
                for (int c = 0; c < C; ++c) v = fmaf(W2[c * Cr + r], ds[c], v);
                dh[i2] = hid[i2] > 0.f ? v : 0.f;
            }
            {   // dW2[c][r] = ds[c] * (h_avg[r] + h_max[r])
                float* d2 = wrow + P.o_up_w2[j];
                for (int i2 = tid; i2 < C * Cr; i2 += ST_NT) {
                    const int c = i2 / Cr, r = i2 - c * Cr;
                    const float v = ds[c] * (hid[r] + hid[Cr + r]);
                    d2[i2] = accw ? d2[i2] + v : v;
                }
            }
            __syncthreads();
            {   // dW1[r][c] = dh_avg[r] * avg[c] + dh_max[r] * max[c]
                float* d1 = wrow + P.o_up_w1[j];
                for (int i2 = tid; i2 < Cr * C; i2 += ST_NT) {
                    const int r = i2 / C, c = i2 - r * C;
                    const float v = fmaf(dh[r], pool[c], dh[Cr + r] * pool[C + c]);
                    d1[i2] = accw ? d1[i2] + v : v;
                }
            }
            for (int c = tid; c < C; c += ST_NT) {
                float ga = 0.f, gm = 0.f;
                for (int r = 0; r < Cr; ++r) {
                    const float w = W1[r * C + c];
                    ga = fmaf(w, dh[r], ga);
                    gm = fmaf(w, dh[Cr + r], gm);
                }
                dav[c] = ga / (float)HW;
                dmx[c] = gm;
            }
            __syncthreads();
            // dy = att * g + davg + dmax * [pixel == argmax]; per-channel sums -> bias gradient of the transposed conv
            float* dy = GA(si, P.g_a);
            for (int c = warp; c < C; c += ST_NW) {
                const float at = att[c], da = dav[c], dm = dmx[c];
                const int amax = (int)pool[2 * C + c];
                float s = 0.f;
                for (int i2 = lane; i2 < HW; i2 += 32) {
                    float v = fmaf(at, dcat[c * HW + i2], da);
                    if (i2 == amax) v += dm;
                    dy[c * HW + i2] = v;
                    s += v;
                }
                const double Sd = warp_sum_d((double)s);
                if (lane == 0) {
                    float* dbp = wrow + P.o_up_b[j] + c;
                    *dbp = accw ? *dbp + (float)Sd : (float)Sd;
                }
            }
            // the block's input activation (recomputed)
            const uint32_t base = (uint32_t)(n0 + si) * (uint32_t)in_elems;
            if (j == 0) {
                const CaeStemTrainFc& F = S.fc[S.n_fc - 1];
                if (F.has_bn) st_make_act(act, TP(si, P.t[S.n_fc - 1]), in_elems, 1, ST_COEF(4 + S.n_fc - 1), ST_COEF(4 + S.n_fc - 1) + ST_CMAX,
                                          true, drop, 4 + S.n_fc - 1, base);
                else st_make_act(act, TP(si, P.t[S.n_fc - 1]), in_elems, 1, nullptr, nullptr, false, drop, 4 + S.n_fc - 1, base);
            } else {
                const CaeStemTrainUp& Lp = S.up[j - 1];
                st_make_act(act, TP(si, P.cat[j - 1]), 2 * Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(8 + j - 1), ST_COEF(8 + j - 1) + ST_CMAX, true,
                            drop, 8 + j - 1, base);
            }
            __syncthreads();
            // weight gradient dW[ci][co][t] = sum_in act[ci][iy][ix] * dy[co][iy*s - p + ky][..]
            st_wgrad(act, L.Hin, L.Win, dy, L.Hout, L.Wout, L.Cin, C, L.k, L.stride, L.pad, wrow + P.o_up_w[j], accw);
            // input gradient (wrt the activated, dropped input)
            float* gin = GA(si, P.g_b);
            st_sconv(dy, C, L.Hout, L.Wout, gin, L.Cin, L.Hin, L.Win, L.k, L.stride, L.pad, wsm, nullptr);
            __syncthreads();
            if (j > 0) {
                const CaeStemTrainUp& Lp = S.up[j - 1];
                const int C2 = 2 * Lp.Cout, HWp = Lp.Hout * Lp.Wout;
                float* dz = GA(si, P.g_dz);
                mask_to_dz(dz, gin, TP(si, P.cat[j - 1]), C2, HWp, ST_COEF(8 + j - 1), 8 + j - 1, base, nullptr);
                __syncthreads();
                st_chan_stats_bwd(dz, TP(si, P.cat[j - 1]), C2, HWp, ST_COEF(8 + j - 1) + 2 * ST_CMAX, ST_COEF(8 + j - 1) + 3 * ST_CMAX, sacc);
                __syncthreads();
            } else {
                // ---- decoder_lin.4 (no BN): u = relu(W a3 + b); gin = dL/d dropout(u)
                const int i4 = S.n_fc - 1, i3 = S.n_fc - 2;
                const CaeStemTrainFc& F4 = S.fc[i4];
                const CaeStemTrainFc& F3 = S.fc[i3];
                float* dpre = GA(si, P.g_a);
                const float* u = TP(si, P.t[i4]);
                for (int e = tid; e < F4.out; e += ST_NT) {
                    float v = gin[e];
                    if (drop.thresh) v = st_keep(drop, 4 + i4, base + e) ? v * drop.scale : 0.f;
                    dpre[e] = u[e] > 0.f ? v : 0.f;
                }
                // its input a3 = dropout(relu(bn1d(t3)))
                const uint32_t base3 = (uint32_t)(n0 + si) * (uint32_t)F4.in;
                st_make_act(act, TP(si, P.t[i3]), F4.in, 1, ST_COEF(4 + i3), ST_COEF(4 + i3) + ST_CMAX, true, drop, 4 + i3, base3);
                __syncthreads();
                st_fc_dw(dpre, F4.out, act, F4.in, wrow + P.o_fc_w[i4], wrow + P.o_fc_b[i4], accw);
                float* da3 = GA(si, P.g_b);
                st_fc_dx(PRM(F4.w), dpre, F4.out, da3, F4.in);
                __syncthreads();
                float* dz = GA(si, P.g_dz);
                mask_to_dz(dz, da3, TP(si, P.t[i3]), F3.out, 1, ST_COEF(4 + i3), 4 + i3, base3, nullptr);
                __syncthreads();
                st_chan_stats_bwd(dz, TP(si, P.t[i3]), F3.out, 1, ST_COEF(4 + i3) + 2 * ST_CMAX, ST_COEF(4 + i3) + 3 * ST_CMAX, sacc);
                __syncthreads();
            }
        }
        if (j > 0) {
            const CaeStemTrainUp& Lp = S.up[j - 1];
            st_grid_sums(grid, S.bnpart, parity, 2 * Lp.Cout, sacc, tot, ti);
            parity ^= 1;
            st_bn_finalize_bwd(Lp.bn, PRM(Lp.bn.gamma), tot, (double)S.N * Lp.Hout * Lp.Wout, ST_COEF(8 + j - 1), bc);
        } else {
            const int i3 = S.n_fc - 2;
            st_grid_sums(grid, S.bnpart, parity, S.fc[i3].out, sacc, tot, ti);
            parity ^= 1;
            st_bn_finalize_bwd(S.fc[i3].bn, PRM(S.fc[i3].bn.gamma), tot, (double)S.N, ST_COEF(4 + i3), bc);
        }
    }
    // ---- decoder_lin.0 (BN) <- latent <- encoder_lin.4 (no BN) <- encoder_lin.0 (BN): up to the BatchNorm1d of fc[0]
    {
        const CaeStemTrainFc& F3 = S.fc[2];
        const CaeStemTrainFc& F2 = S.fc[1];
        const CaeStemTrainFc& F1 = S.fc[0];
        st_zero_acc(sacc, 2 * F1.out);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const bool accw = si > 0;
            float* dt3 = GA(si, P.g_dz);
            bn_apply_bwd(dt3, TP(si, P.t[2]), F3.out, 1);
            // input of fc[2]: dropout(z), z stored post-ReLU
            const uint32_t basez = (uint32_t)(n0 + si) * (uint32_t)F3.in;
            st_make_act(act, TP(si, P.t[1]), F3.in, 1, nullptr, nullptr, false, drop, 4 + 1, basez);
            __syncthreads();
            st_fc_dw(dt3, F3.out, act, F3.in, wrow + P.o_fc_w[2], nullptr, accw);        // bias in front of a BatchNorm: dead
            float* dzd = GA(si, P.g_a);
            st_fc_dx(PRM(F3.w), dt3, F3.out, dzd, F3.in);
            __syncthreads();
            float* dpre2 = GA(si, P.g_b);
            const float* z = TP(si, P.t[1]);
            for (int e = tid; e < F2.out; e += ST_NT) {
                float v = dzd[e];
                if (drop.thresh) v = st_keep(drop, 4 + 1, basez + e) ? v * drop.scale : 0.f;
                dpre2[e] = z[e] > 0.f ? v : 0.f;
            }
            const uint32_t base1 = (uint32_t)(n0 + si) * (uint32_t)F2.in;
            st_make_act(act, TP(si, P.t[0]), F2.in, 1, ST_COEF(4 + 0), ST_COEF(4 + 0) + ST_CMAX, true, drop, 4 + 0, base1);
            __syncthreads();
            st_fc_dw(dpre2, F2.out, act, F2.in, wrow + P.o_fc_w[1], wrow + P.o_fc_b[1], accw);
            float* da1 = GA(si, P.g_a);
            st_fc_dx(PRM(F2.w), dpre2, F2.out, da1, F2.in);
            __syncthreads();
            float* dz1 = GA(si, P.g_dz);
            mask_to_dz(dz1, da1, TP(si, P.t[0]), F1.out, 1, ST_COEF(4 + 0), 4 + 0, base1, nullptr);
            __syncthreads();
            st_chan_stats_bwd(dz1, TP(si, P.t[0]), F1.out, 1, ST_COEF(4 + 0) + 2 * ST_CMAX, ST_COEF(4 + 0) + 3 * ST_CMAX, sacc);
            __syncthreads();
        }
        st_grid_sums(grid, S.bnpart, parity, F1.out, sacc, tot, ti);
        parity ^= 1;
        st_bn_finalize_bwd(F1.bn, PRM(F1.bn.gamma), tot, (double)S.N, ST_COEF(4 + 0), bc);
    }
    // ---- encoder_lin.0 -> last encoder conv's BatchNorm
    {
        const CaeStemTrainFc& F1 = S.fc[0];
        const int le = S.n_conv - 1;
        const CaeStemTrainConv& Le = S.conv[le];
        const int HWe = Le.Hout * Le.Wout;
        st_zero_acc(sacc, 2 * Le.Cout);
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const bool accw = si > 0;
            float* dt1 = GA(si, P.g_dz);
            bn_apply_bwd(dt1, TP(si, P.t[0]), F1.out, 1);
            const uint32_t base = (uint32_t)(n0 + si) * (uint32_t)F1.in;
            st_make_act(act, TP(si, P.ye[le]), Le.Cout, HWe, ST_COEF(le), ST_COEF(le) + ST_CMAX, true, drop, le, base);
            __syncthreads();
            st_fc_dw(dt1, F1.out, act, F1.in, wrow + P.o_fc_w[0], nullptr, accw);
            float* da = GA(si, P.g_a);
            st_fc_dx(PRM(F1.w), dt1, F1.out, da, F1.in);
            __syncthreads();
            float* dz = GA(si, P.g_dz);
            mask_to_dz(dz, da, TP(si, P.ye[le]), Le.Cout, HWe, ST_COEF(le), le, base, nullptr);
            __syncthreads();
            st_chan_stats_bwd(dz, TP(si, P.ye[le]), Le.Cout, HWe, ST_COEF(le) + 2 * ST_CMAX, ST_COEF(le) + 3 * ST_CMAX, sacc);
            __syncthreads();
        }
        st_grid_sums(grid, S.bnpart, parity, Le.Cout, sacc, tot, ti);
        parity ^= 1;
        st_bn_finalize_bwd(Le.bn, PRM(Le.bn.gamma), tot, (double)S.N * HWe, ST_COEF(le), bc);
    }
    // ---- encoder convs, last to first
    for (int l = S.n_conv - 1; l >= 0; --l) {
        const CaeStemTrainConv& L = S.conv[l];
        const int HWo = L.Hout * L.Wout, in_elems = L.Cin * L.Hin * L.Win, KK = L.k * L.k;
        if (l > 0) {
            st_stage_w(wsm, PRM(L.w), L.Cout, L.Cin, KK, true);   // input gradient: transposed gather over dy with w[co][t][ci]
            st_zero_acc(sacc, 2 * S.conv[l - 1].Cout);
        }
        __syncthreads();
        for (int si = 0; si < ns; ++si) {
            const bool accw = si > 0;
            float* dy = GA(si, P.g_dz);
            bn_apply_bwd(dy, TP(si, P.ye[l]), L.Cout, HWo);
            const float* in = TP(si, P.x);
            const uint32_t base = (uint32_t)(n0 + si) * (uint32_t)in_elems;
            if (l > 0) {
                const CaeStemTrainConv& Lp = S.conv[l - 1];
                st_make_act(act, TP(si, P.ye[l - 1]), Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(l - 1), ST_COEF(l - 1) + ST_CMAX, true, drop,
                            l - 1, base);
                in = act;
            }
            __syncthreads();
            // dW[co][ci][t] = sum_out dy[co][oy][ox] * in[ci][oy*s - p + ky][..]
            st_wgrad(dy, L.Hout, L.Wout, in, L.Hin, L.Win, L.Cout, L.Cin, L.k, L.stride, L.pad, wrow + P.o_conv_w[l], accw);
            if (l > 0) {
                const CaeStemTrainConv& Lp = S.conv[l - 1];
                float* gin = GA(si, P.g_a);
                st_tconv(dy, L.Cout, L.Hout, L.Wout, gin, L.Cin, L.Hin, L.Win, L.k, L.stride, L.pad, wsm, nullptr);
                __syncthreads();
                // add the skip-connection gradient (if a decoder block read this activation) after the dropout mask
                bool has_skip = false;
                for (int j = 0; j < S.n_up; ++j) has_skip = has_skip || (S.up[j].skip == l - 1);
                float* dzn = GA(si, P.g_dz);       // dy (same buffer) was last read before the barrier above
                mask_to_dz(dzn, gin, TP(si, P.ye[l - 1]), Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(l - 1), l - 1, base,
                           has_skip ? GA(si, P.g_skip[l - 1]) : nullptr);
                __syncthreads();
                st_chan_stats_bwd(dzn, TP(si, P.ye[l - 1]), Lp.Cout, Lp.Hout * Lp.Wout, ST_COEF(l - 1) + 2 * ST_CMAX,
                                  ST_COEF(l - 1) + 3 * ST_CMAX, sacc);
            }
            __syncthreads();
        }
        if (l > 0) {
            const CaeStemTrainConv& Lp = S.conv[l - 1];
            st_grid_sums(grid, S.bnpart, parity, Lp.Cout, sacc, tot, ti);
            parity ^= 1;
            st_bn_finalize_bwd(Lp.bn, PRM(Lp.bn.gamma), tot, (double)S.N * Lp.Hout * Lp.Wout, ST_COEF(l - 1), bc);
        }
    }
    // ---- weight gradients: fixed-order sum of the per-CTA rows, spread over the grid
    ST_T(ti++);
    __threadfence();
    grid.sync();
    ST_T(ti++);
    {
        const int rows = gridDim.x;
        for (int e = blockIdx.x * ST_NT + tid; e < P.wrow; e += gridDim.x * ST_NT) {
            float s = 0.f;
            {
                const float* cp = S.wpart + e;
                int r = 0;
                for (; r + 16 <= rows; r += 16) {
                    float v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) v[u] = __ldcg(cp + (size_t)(r + u) * P.wrow);
#pragma unroll
                    for (int u = 0; u < 16; ++u) s += v[u];
                }
                for (; r < rows; ++r) s += __ldcg(cp + (size_t)r * P.wrow);
            }
            // scatter to the parameter's gradient tensor
            float* dst = nullptr;
            int off = 0;
            for (int l = 0; l < S.n_conv && !dst; ++l) {
                const int n = S.conv[l].Cout * S.conv[l].Cin * S.conv[l].k * S.conv[l].k;
                if (e >= P.o_conv_w[l] && e < P.o_conv_w[l] + n) { dst = S.conv[l].dw; off = e - P.o_conv_w[l]; }
            }
            for (int i = 0; i < S.n_fc && !dst; ++i) {
                const int n = S.fc[i].in * S.fc[i].out;
                if (e >= P.o_fc_w[i] && e < P.o_fc_w[i] + n) { dst = S.fc[i].dw; off = e - P.o_fc_w[i]; }
                else if (!S.fc[i].has_bn && e >= P.o_fc_b[i] && e < P.o_fc_b[i] + S.fc[i].out) { dst = S.fc[i].db; off = e - P.o_fc_b[i]; }
            }
            for (int j = 0; j < S.n_up && !dst; ++j) {
                const CaeStemTrainUp& U = S.up[j];
                const int n = U.Cin * U.Cout * U.k * U.k, n1 = U.Cr * U.Cout;
                if (e >= P.o_up_w[j] && e < P.o_up_w[j] + n) { dst = U.dw; off = e - P.o_up_w[j]; }
                else if (e >= P.o_up_b[j] && e < P.o_up_b[j] + U.Cout) { dst = U.db; off = e - P.o_up_b[j]; }
                else if (e >= P.o_up_w1[j] && e < P.o_up_w1[j] + n1) { dst = U.dW1; off = e - P.o_up_w1[j]; }
                else if (e >= P.o_up_w2[j] && e < P.o_up_w2[j] + n1) { dst = U.dW2; off = e - P.o_up_w2[j]; }
            }
            if (dst) dst[off] = s;
        }
    }
    ST_T(ti++);
    if (blockIdx.x == 0 && threadIdx.x == 0) st_prof[63] = (unsigned long long)ti;
#undef TP
#undef GA
#undef PRM
}

// ============================================================================================================
// host
// ============================================================================================================
static int st_plan(const CaeStemTrain& s, StPlan& p, char* why, size_t why_len) {
#define ST_FAIL(...) do { snprintf(why, why_len, __VA_ARGS__); return 0; } while (0)
    memset(&p, 0, sizeof(p));
    if (s.n_conv < 1 || s.n_conv > CAE_STEM_MAX || s.n_up < 1 || s.n_up > CAE_STEM_MAX) ST_FAIL("layer counts");
    if (s.n_fc != 4 || !s.fc[0].has_bn || s.fc[1].has_bn || !s.fc[2].has_bn || s.fc[3].has_bn)
        ST_FAIL("fc stacks must be Linear-BN-ReLU-Linear-ReLU twice (unet.py:92-100,121-129)");
    if (s.N < 1) ST_FAIL("empty batch");
    p.spc = (s.N + CAE_NUM_SMS - 1) / CAE_NUM_SMS;
    if (p.spc > ST_MAX_SPC) ST_FAIL("batch %d needs %d samples per CTA (> %d)", s.N, p.spc, ST_MAX_SPC);
    p.ctas = (s.N + p.spc - 1) / p.spc;
    int off = 0, actmax = 0, wmax = 0;
    auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
    auto up4 = [](int v) { return (v + 3) & ~3; };
    int prev = s.conv[0].Cin * s.conv[0].Hin * s.conv[0].Win;
    p.x = take(prev);
    actmax = prev;
    for (int l = 0; l < s.n_conv; ++l) {
        const CaeStemTrainConv& L = s.conv[l];
        if (!L.w || !L.dw || L.Cin * L.Hin * L.Win != prev) ST_FAIL("encoder layer %d does not chain", l);
        if ((L.Hin + 2 * L.pad - L.k) / L.stride + 1 != L.Hout || (L.Win + 2 * L.pad - L.k) / L.stride + 1 != L.Wout)
            ST_FAIL("encoder layer %d geometry", l);
        if (L.Cout > ST_CMAX || L.bn.C != L.Cout || !L.bn.scale || !L.bn.shift || !L.bn.mean || !L.bn.invstd) ST_FAIL("encoder layer %d BN", l);
        prev = L.Cout * L.Hout * L.Wout;
        p.ye[l] = take(prev);
        actmax = max(actmax, prev);
        wmax = max(wmax, max(L.Cin * L.k * L.k * up4(L.Cout), L.Cout * L.k * L.k * up4(L.Cin)));
    }
    for (int i = 0; i < s.n_fc; ++i) {
        const CaeStemTrainFc& L = s.fc[i];
        if (!L.w || !L.dw || L.in != prev) ST_FAIL("fc layer %d does not chain (%d vs %d)", i, L.in, prev);
        if (L.has_bn && (L.out > ST_CMAX || L.bn.C != L.out || !L.bn.scale || !L.bn.mean)) ST_FAIL("fc layer %d BN (<= %d features)", i, ST_CMAX);
        if (!L.has_bn && !L.db) ST_FAIL("fc layer %d needs a bias gradient", i);
        prev = L.out;
        p.t[i] = take(prev);
        actmax = max(actmax, prev);
    }
    for (int j = 0; j < s.n_up; ++j) {
        const CaeStemTrainUp& L = s.up[j];
        if (!L.w || !L.dw || !L.W1 || !L.W2 || !L.dW1 || !L.dW2 || !L.db || L.Cin * L.Hin * L.Win != prev) ST_FAIL("decoder block %d does not chain", j);
        if ((L.Hin - 1) * L.stride - 2 * L.pad + L.k != L.Hout || (L.Win - 1) * L.stride - 2 * L.pad + L.k != L.Wout)
            ST_FAIL("decoder block %d geometry (output padding is not supported)", j);
        if (L.skip < 0 || L.skip >= s.n_conv) ST_FAIL("decoder block %d: bad skip index", j);
        const CaeStemTrainConv& E = s.conv[L.skip];
        if (E.Cout != L.Cout || E.Hout != L.Hout || E.Wout != L.Wout) ST_FAIL("decoder block %d: skip geometry mismatch", j);
        if (2 * L.Cout > ST_CMAX || L.bn.C != 2 * L.Cout || !L.bn.scale || !L.bn.mean || L.Cr < 1 || 2 * L.Cr > ST_CMAX)
            ST_FAIL("decoder block %d BN / attention size", j);
        const int yel = L.Cout * L.Hout * L.Wout;
        p.yu[j] = take(yel);
        p.cat[j] = take(2 * yel);
        p.att[j] = take(L.Cout);
        p.hid[j] = take(2 * L.Cr);
        p.pool[j] = take(3 * L.Cout);
        prev = 2 * yel;
        actmax = max(actmax, prev);
        wmax = max(wmax, max(L.Cin * L.k * L.k * up4(L.Cout), L.Cout * L.k * L.k * up4(L.Cin)));
    }
    p.tape = off;
    p.act_max = up4(actmax);
    p.w_max = up4(wmax);
    // weight-gradient row
    int wo = 0;
    auto wtake = [&](int n) { int o = wo; wo += (n + 3) & ~3; return o; };
    for (int l = 0; l < s.n_conv; ++l) p.o_conv_w[l] = wtake(s.conv[l].Cout * s.conv[l].Cin * s.conv[l].k * s.conv[l].k);
    for (int i = 0; i < s.n_fc; ++i) { p.o_fc_w[i] = wtake(s.fc[i].in * s.fc[i].out); p.o_fc_b[i] = wtake(s.fc[i].out); }
    for (int j = 0; j < s.n_up; ++j) {
        p.o_up_w[j] = wtake(s.up[j].Cin * s.up[j].Cout * s.up[j].k * s.up[j].k);
        p.o_up_b[j] = wtake(s.up[j].Cout);
        p.o_up_w1[j] = wtake(s.up[j].Cr * s.up[j].Cout);
        p.o_up_w2[j] = wtake(s.up[j].Cr * s.up[j].Cout);
    }
    p.wrow = wo;
    // gradient arena
    int go = 0;
    auto gtake = [&](int n) { int o = go; go += (n + 3) & ~3; return o; };
    p.g_dz = gtake(p.act_max);
    p.g_a = gtake(p.act_max);
    p.g_b = gtake(p.act_max);
    for (int l = 0; l < s.n_conv; ++l) p.g_skip[l] = gtake(s.conv[l].Cout * s.conv[l].Hout * s.conv[l].Wout);
    p.garena = go;
    // every parameter pointer must lie inside the contiguous block [params, params + params_len)
    if (!s.params || s.params_len < 4 || s.params_len % 4 != 0 || (reinterpret_cast<uintptr_t>(s.params) & 15))
        ST_FAIL("parameter block missing or not 16-byte aligned / sized");
    auto inside = [&](const float* q, long long n) { return q == nullptr || (q >= s.params && q + n <= s.params + s.params_len); };
    bool ok = true;
    for (int l = 0; l < s.n_conv; ++l) {
        const CaeStemTrainConv& L = s.conv[l];
        ok = ok && inside(L.w, (long long)L.Cout * L.Cin * L.k * L.k) && inside(L.b, L.Cout) && inside(L.bn.gamma, L.Cout) && inside(L.bn.beta, L.Cout);
    }
    for (int i = 0; i < s.n_fc; ++i) {
        const CaeStemTrainFc& L = s.fc[i];
        ok = ok && inside(L.w, (long long)L.in * L.out) && inside(L.b, L.out);
        if (L.has_bn) ok = ok && inside(L.bn.gamma, L.out) && inside(L.bn.beta, L.out);
    }
    for (int j = 0; j < s.n_up; ++j) {
        const CaeStemTrainUp& L = s.up[j];
        ok = ok && inside(L.w, (long long)L.Cin * L.Cout * L.k * L.k) && inside(L.b, L.Cout) && inside(L.W1, L.Cr * L.Cout) &&
             inside(L.W2, L.Cr * L.Cout) && inside(L.bn.gamma, 2 * L.Cout) && inside(L.bn.beta, 2 * L.Cout);
    }
    if (!ok) ST_FAIL("a parameter lies outside the contiguous parameter block");
    p.pfloats = (int)s.params_len;
    const size_t common = (size_t)(p.act_max + p.w_max + ST_NBN * 4 * ST_CMAX + p.pfloats) * 4 + (size_t)4 * ST_CMAX * 8 + 16;
    p.smem_fwd = (int)((size_t)p.spc * p.tape * 4 + common);
    p.smem_bwd = (int)((size_t)p.spc * (p.tape + p.garena) * 4 + common + (size_t)(3 + 6) * ST_CMAX * 4);
    if (p.smem_fwd > 220 * 1024 || p.smem_bwd > 220 * 1024)
        ST_FAIL("%d samples per CTA need %d / %d KB of shared memory (> 220)", p.spc, p.smem_fwd / 1024, p.smem_bwd / 1024);
    return 1;
#undef ST_FAIL
}

extern "C" int cae_unet_stem_train_supported(const CaeStemTrain* s) {
    if (!s) return 0;
    StPlan p;
    char why[160];
    return st_plan(*s, p, why, sizeof(why));
}

extern "C" long long cae_unet_stem_train_tape_elems(const CaeStemTrain* s) {
    StPlan p;
    char why[160];
    if (!s || !st_plan(*s, p, why, sizeof(why))) return -1;
    return p.tape;
}

extern "C" long long cae_unet_stem_train_workspace(const CaeStemTrain* s, int which) {
    StPlan p;
    char why[160];
    if (!s || !st_plan(*s, p, why, sizeof(why))) return -1;
    if (which == 0) return 2ll * p.ctas * 2 * ST_CMAX;          // doubles: two parities of [ctas][2*CMAX]
    return (long long)p.ctas * p.wrow;                          // floats
}

template <typename K>
static int st_launch(K kernel, const CaeStemTrain* s, const CaeSrc* x, bool backward, void* stream, const char* what) {
    CAE_REQUIRE(s && x, "%s: null argument", what);
    int rc;
    if ((rc = check_view(x->t0, what))) return rc;
    StArgs a;
    memset(&a, 0, sizeof(a));
    char why[160] = "";
    if (!st_plan(*s, a.p, why, sizeof(why))) {
        cae_set_error("%s: %s", what, why);
        return CAE_EUNSUPPORTED;
    }
    const CaeStemTrainConv& c0 = s->conv[0];
    CAE_REQUIRE(x->t0.N == s->N && x->t0.C == c0.Cin && x->t0.H == c0.Hin && x->t0.W == c0.Win, "%s: input view differs from the first layer", what);
    CAE_REQUIRE(s->tape && s->hin && s->bnpart && s->wpart, "%s: tape / hin / workspaces missing", what);
    CAE_REQUIRE(!backward || s->dhin, "%s: dhin missing", what);
    CAE_REQUIRE(s->dropout_p >= 0.f && s->dropout_p < 1.f, "%s: dropout_p %f outside [0, 1)", what, (double)s->dropout_p);
    a.s = *s;
    a.x = *x;
    const int smem = backward ? a.p.smem_bwd : a.p.smem_fwd;
    ensure_smem_limit(kernel, 220 * 1024);
    void* params[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)kernel, dim3(a.p.ctas), dim3(ST_NT), params, (size_t)smem, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        cae_set_error("%s: cooperative launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return cae_check_launch(what);
}

extern "C" int cae_unet_stem_train_fwd(const CaeStemTrain* s, const CaeSrc* x, void* stream) {
    return st_launch(k_unet_stem_train_fwd, s, x, false, stream, "cae_unet_stem_train_fwd");
}

extern "C" int cae_unet_stem_train_bwd(const CaeStemTrain* s, const CaeSrc* x, void* stream) {
    return st_launch(k_unet_stem_train_bwd, s, x, true, stream, "cae_unet_stem_train_bwd");
}

// phase timestamps (clock64 of CTA 0) of the last forward ([0..31]) and backward ([32..63]) launch; profiling aid
extern "C" int cae_unet_stem_train_profile(unsigned long long* out64) {
    CAE_REQUIRE(out64, "unet_stem_train_profile: null argument");
    cudaError_t e = cudaMemcpyFromSymbol(out64, st_prof, sizeof(unsigned long long) * 64);
    if (e != cudaSuccess) {
        cae_set_error("unet_stem_train_profile: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return CAE_OK;
}
