// cae_gemm on the tensor cores: the nn.Linear contractions of the large-fc regimes (UNET with fc 3200 / latent 800 at batch
// 256: unet.py:92-100,121-129; LinearModel at large batches: linear.py:43) are dense GEMMs with both free dimensions in the
// hundreds - the 32 x 32 fp32 SIMT tile of k_gemm reaches ~10 TFLOP/s there.  This route keeps cae_gemm's contract
// (strided operands, on-load BatchNorm + ReLU affine, bias / ReLU / mask epilogue, bias-gradient row sums) and runs the
// contraction through cae_tc_gemm (tcgen05 kind::tf32, 3xTF32):
//   k_td_split     operand -> (hi, lo) in its own major (K-major or MN-major, pitch padded to 4 floats), affine applied
//   cae_tc_gemm    split-K partial tiles into the workspace
//   k_td_epilogue  fixed-order sum of the K slices + bias, ReLU, mask -> C
//   k_td_rowsum    rowsum_A (bias gradient), fixed order
// Eligibility (cae_gemm_tc_workspace > 0): min(M, N) >= 128, so that the two split passes (M*K + N*K elements) stay small
// against the M*N*K contraction; K >= 32; >= 0.1 GFLOP; every operand contiguous along one of its two axes; C row-major.
// A 256 -> 65 536 Linear at batch 64 is NOT eligible on purpose: it is bound by one read of the 67 MB weight, which the
// split would triple.
#include "capi_host.h"
#include "common.cuh"

struct TdPlan {
    int a_kmajor, b_kmajor;
    long long rowsA, lda, rowsB, ldb, ldp, split_stride;
    int splits, tile_n;
    long long ws;
};

static bool td_plan(const CaeGemm* g, TdPlan& p) {
    if (!g || g->M < 128 || g->N < 128 || g->K < 32) return false;
    if ((double)g->M * g->N * g->K < 5e7) return false;
    if (!(g->sAk == 1 || g->sAm == 1) || !(g->sBk == 1 || g->sBn == 1) || g->sCn != 1) return false;
    p.a_kmajor = g->sAk == 1;                 // A(m,k) = Ap[m*sAm + k]      else Ap[m + k*sAk]
    p.b_kmajor = g->sBk == 1;                 // B(k,n) = Bp[k + n*sBn]      else Bp[k*sBk + n]
    p.rowsA = p.a_kmajor ? g->M : g->K; p.lda = roundup4(p.a_kmajor ? g->K : g->M);
    p.rowsB = p.b_kmajor ? g->N : g->K; p.ldb = roundup4(p.b_kmajor ? g->K : g->N);
    p.tile_n = g->N >= 256 ? 256 : 128;
    const long long tiles = (long long)ceil_div(g->M, 128) * ceil_div(g->N, p.tile_n);
    const int nkb = ceil_div(g->K, 32);
    long long want = tiles >= CAE_NUM_SMS ? 1 : ceil_div(CAE_NUM_SMS, tiles);
    const long long cap = nkb / 4 > 1 ? nkb / 4 : 1;            // at least four K blocks per slice
    if (want > cap) want = cap;
    const long long acc = ceil_div(nkb, 64);                    // at most 64 K blocks per slice (fp32 accumulation in the MMA)
    if (want < acc) want = acc;
    const int kbps = ceil_div(nkb, (int)want);
    p.splits = ceil_div(nkb, kbps);
    p.ldp = roundup4(g->N);
    p.split_stride = (long long)g->M * p.ldp;
    p.ws = 2 * p.rowsA * p.lda + 2 * p.rowsB * p.ldb + (long long)p.splits * p.split_stride;
    return true;
}

// element (o, i) of the operand's own major: inner index i is the contiguous one.  k_major: (row = o, k = i) else (row = i, k = o)
__global__ void __launch_bounds__(CAE_NT) k_td_split(const float* __restrict__ X, long long s_outer, long long outer, long long inner,
                                                     const float* __restrict__ k0, const float* __restrict__ k2, int hw, int relu,
                                                     int chan_is_inner, float* __restrict__ hi, float* __restrict__ lo, long long ld) {
    const long long total = outer * inner;
    for (long long e = (long long)blockIdx.x * CAE_NT + threadIdx.x; e < total; e += (long long)gridDim.x * CAE_NT) {
        const long long o = e / inner, i = e - o * inner;
        float v = __ldg(X + o * s_outer + i);
        if (k0) {
            const long long c = (chan_is_inner ? i : o) / hw;
            v = fmaf(v, __ldg(k0 + c), __ldg(k2 + c));
        }
        if (relu) v = fmaxf(v, 0.f);
        float h, l;
        tf32_split(v, h, l);
        hi[o * ld + i] = h;
        lo[o * ld + i] = l;
    }
}

__global__ void __launch_bounds__(CAE_NT) k_td_epilogue(const float* __restrict__ part, int splits, long long split_stride, int M, int N,
                                                        long long ldp, float* __restrict__ Cp, long long sCm,
                                                        const float* __restrict__ bias, int relu, const float* __restrict__ mask) {
    const long long total = (long long)M * N;
    for (long long e = (long long)blockIdx.x * CAE_NT + threadIdx.x; e < total; e += (long long)gridDim.x * CAE_NT) {
        const long long m = e / N, n = e - m * N;
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += __ldg(part + z * split_stride + m * ldp + n);
        if (bias) v += __ldg(bias + n);
        if (relu) v = fmaxf(v, 0.f);
        if (mask && !(__ldg(mask + m * sCm + n) > 0.f)) v = 0.f;
        Cp[m * sCm + n] = v;
    }
}

// rowsum_A[m] = sum_k A(m,k), fixed order.  A contiguous along m: CTA = 32 adjacent m x 8 k lanes (coalesced 128-byte rows,
// four independent loads in flight per thread - a thread-per-m loop over k paid one L2 round trip per k: 60 us at K = 256);
// otherwise one warp per m.
__global__ void __launch_bounds__(CAE_NT) k_td_rowsum(const CaeGemm g) {
    auto val = [&](long long m, long long k) {
        float v = __ldg(g.A + m * g.sAm + k * g.sAk);
        if (g.a_k0) v = fmaf(v, __ldg(g.a_k0 + k / g.a_hw), __ldg(g.a_k2 + k / g.a_hw));
        return g.a_relu ? fmaxf(v, 0.f) : v;
    };
    if (g.sAm == 1) {
        __shared__ float s_p[8][32];
        const int ml = threadIdx.x & 31, kl = threadIdx.x >> 5;
        const long long m = (long long)blockIdx.x * 32 + ml;
        float s = 0.f;
        if (m < g.M) {
            int k = kl;
            for (; k + 24 < g.K; k += 32) {
                const float v0 = val(m, k), v1 = val(m, k + 8), v2 = val(m, k + 16), v3 = val(m, k + 24);
                s += v0; s += v1; s += v2; s += v3;
            }
            for (; k < g.K; k += 8) s += val(m, k);
        }
        s_p[kl][ml] = s;
        __syncthreads();
        if (kl == 0 && m < g.M) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += s_p[q][ml];
            g.rowsum_A[m] = t;
        }
    } else {
        const long long m = (long long)blockIdx.x * CAE_NWARP + (threadIdx.x >> 5);
        if (m < g.M) {
            float s = 0.f;
            for (int k = threadIdx.x & 31; k < g.K; k += 32) s += val(m, k);
            s = warp_sum(s);
            if ((threadIdx.x & 31) == 0) g.rowsum_A[m] = s;
        }
    }
}

extern "C" long long cae_gemm_tc_workspace(const CaeGemm* g) {
    TdPlan p;
    return td_plan(g, p) ? p.ws : 0;
}

extern "C" int cae_gemm_tc(const CaeGemm* g, float* ws, long long ws_len, void* stream) {
    CAE_REQUIRE(g && g->A && g->B && g->C && ws, "gemm_tc: null argument");
    CAE_REQUIRE((!g->a_k0 || (g->a_k2 && g->a_hw > 0)) && (!g->b_k0 || (g->b_k2 && g->b_hw > 0)),
                "gemm_tc: on-load affine needs k0, k2 and hw");
    TdPlan p;
    if (!td_plan(g, p)) {
        cae_set_error("gemm_tc: %dx%dx%d with these strides is not eligible (cae_gemm_tc_workspace == 0): use cae_gemm", g->M, g->N, g->K);
        return CAE_EUNSUPPORTED;
    }
    CAE_REQUIRE(ws_len >= p.ws && (uintptr_t)ws % 16 == 0, "gemm_tc: workspace of %lld floats (16-byte aligned) needed, got %lld", p.ws,
                ws_len);
    cudaStream_t st = (cudaStream_t)stream;
    float* a_hi = ws;
    float* a_lo = a_hi + p.rowsA * p.lda;
    float* b_hi = a_lo + p.rowsA * p.lda;
    float* b_lo = b_hi + p.rowsB * p.ldb;
    float* part = b_lo + p.rowsB * p.ldb;
    auto blocks = [](long long n) { return (int)min((long long)CAE_NUM_SMS * 8, (n + CAE_NT - 1) / CAE_NT); };
    {   // A: channel index is k
        const long long outer = p.rowsA, inner = p.a_kmajor ? g->K : g->M;
        k_td_split<<<blocks(outer * inner), CAE_NT, 0, st>>>(g->A, p.a_kmajor ? g->sAm : g->sAk, outer, inner, g->a_k0, g->a_k2,
                                                             g->a_hw > 0 ? g->a_hw : 1, g->a_relu, p.a_kmajor ? 1 : 0, a_hi, a_lo, p.lda);
    }
    {   // B: channel index is n
        const long long outer = p.rowsB, inner = p.b_kmajor ? g->K : g->N;
        k_td_split<<<blocks(outer * inner), CAE_NT, 0, st>>>(g->B, p.b_kmajor ? g->sBn : g->sBk, outer, inner, g->b_k0, g->b_k2,
                                                             g->b_hw > 0 ? g->b_hw : 1, g->b_relu, p.b_kmajor ? 0 : 1, b_hi, b_lo, p.ldb);
    }
    int rc = cae_check_launch("cae_gemm_tc(split)");
    if (rc) return rc;
    CaeTcGemm t{};
    t.M = g->M; t.N = g->N; t.K = g->K;
    t.a_hi = a_hi; t.a_lo = a_lo; t.lda = p.lda; t.a_mn_major = !p.a_kmajor;
    t.b_hi = b_hi; t.b_lo = b_lo; t.ldb = p.ldb; t.b_mn_major = !p.b_kmajor;
    t.C = part; t.ldc = p.ldp; t.splits = p.splits; t.split_stride = p.split_stride; t.tile_n = p.tile_n;
    if ((rc = cae_tc_gemm(&t, stream))) return rc;
    k_td_epilogue<<<blocks((long long)g->M * g->N), CAE_NT, 0, st>>>(part, p.splits, p.split_stride, g->M, g->N, p.ldp, g->C, g->sCm,
                                                                     g->bias, g->relu_out, g->mask);
    if (g->rowsum_A) {
        const int grid = g->sAm == 1 ? ceil_div(g->M, 32) : ceil_div(g->M, CAE_NWARP);
        k_td_rowsum<<<grid, CAE_NT, 0, st>>>(*g);
    }
    return cae_check_launch("cae_gemm_tc");
}
