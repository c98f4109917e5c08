// tcgen05 / TMEM / TMA GEMM for the genuinely dense contractions of the path (sm_100a only).
//
//   C[m, n] = sum_k A[m, k] * B[n, k]           fp32 in, fp32 out, fp32 accumulation in tensor memory
//
// Precision: the tensor cores multiply TF32 (10-bit mantissa).  To keep the path's 1e-4 parity bar the operands arrive
// PRE-SPLIT (cae_tc_split / the conv-specific producers in tc_conv.cu): x = hi + lo, hi = x ROUNDED to TF32, lo = x - hi
// rounded to TF32 (common.cuh: tf32_split; what the split drops is <= 2^-22 |x|, unbiased).  Three MMAs per K step -
// lo*hi + hi*lo + hi*hi ("3xTF32"; the dropped lo*lo is <= 2^-24 of the product).  lo == NULL for both operands selects plain
// 1xTF32 (one MMA; ~1e-3 relative).
// Accumulation: the tensor core TRUNCATES its fp32 accumulator once per MMA.  Left to run over the whole K range that
// is a bias growing linearly with K (3e-6 of the max-norm at K = 1000, ~10x an fp32 FMA chain) - harmless per element, but
// BASELINE configs[3] at batch 128 amplifies rounding noise ~2000x on its way back through the decoder (torch's own
// fp32 gradients sit 2e-4 from float64 there) and the tensor-core layers' gradients ended up 4.5e-3 away.  The TMEM
// accumulator therefore only ever holds a CHUNK of KB_PER_CHUNK K blocks: two TMEM accumulators alternate, and while the
// MMA issuer fills one, the epilogue warps drain the other into fp32 REGISTER accumulators with round-to-nearest adds
// (the "promotion" scheme of fp8 GEMMs).  Truncation then happens at the magnitude of a 64-element partial sum only.
//
// Operands may be K-major (K contiguous: A[m*lda + k]) or MN-major (M or N contiguous: A[k*lda + m]); the weight
// gradient of a transposed convolution contracts over positions, which is the LEADING dimension of both of its
// operands as the forward pass lays them out, so it runs MN-major on the same buffers (no transposed copies).
//
// Structure (one 128 x BN output tile per CTA, BK = 32 floats = one 128-byte swizzle row):
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor.2d -> 128B-swizzled shared-memory stages, mbarrier expect_tx
//   warp 1 (one lane)  MMA issuer:   tcgen05.mma.cta_group::1.kind::tf32, accumulator in TMEM, tcgen05.commit -> mbarriers
//   warps 2-9          epilogue:     per chunk tcgen05.ld (32 lanes x 32 columns per instruction) -> += registers; -> global
//                                    (warp w owns TMEM lanes [32 (w % 4), +32) and one half of the BN columns)
// Split-K (gridDim.z) writes partial tiles to C + z * split_stride; the caller reduces them in a fixed order.
//
// Replaces: the cuBLAS/cuDNN GEMMs behind torch.nn.ConvTranspose2d forward/backward for the fat decoder layers
// (reference decoder.py:44-48 executed through aten::convolution / convolution_backward) and torch.nn.Linear
// (linear.py:43, unet.py:92-100,121-129).
#include <cuda.h>
#include "capi_host.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                      // floats per K block = 128 bytes = one swizzle row
constexpr int TILE_A_BYTES = BM * BK * 4;   // 16 KB
constexpr int NUM_THREADS = 320;            // TMA warp, MMA warp, 8 epilogue warps
constexpr int KB_PER_CHUNK = 2;             // K blocks accumulated in TMEM before promotion to registers

struct TcParams {
    int M, N, K;
    int a_mn, b_mn;            // operand is MN-major
    int three_pass;            // 3xTF32 (lo operands present)
    int kb_per_split;          // K blocks per gridDim.z slice
    float* C;
    long long ldc, split_stride;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate; issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start address >> 4 in [0,14), leading byte
// offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type in [61,64) (2 = 128B swizzle)
// (2 = 128B swizzle of 16-byte chunks; 1 = 128B swizzle of 32-byte chunks, the only layout MN-major TF32 operands have)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

#define TC_LD32(taddr, v)                                                                                              \
    asm volatile(                                                                                                      \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                      \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),       \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),      \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                    \
        : "r"(taddr)                                                                                                   \
        : "memory")

// One K block of one operand into its stage slot.
//   K-major : one box {32 k, ROWS rows}  -> [ROWS][128 B], rows swizzled in groups of 8 (1 KB atoms along M/N)
//   MN-major: ROWS/32 boxes {32 mn, 32 k} -> [ROWS/32][32 k-rows][128 B], 32-byte chunks swizzled over 4 rows (512 B
//             atoms along K, 4 KB between MN blocks): cute's Layout_MN_SW128_32B_Atom, the one MN-major TF32 layout
template <int ROWS>
__device__ __forceinline__ void load_operand(uint32_t dst, const CUtensorMap* map, int mn_major, int row0, int k0, uint32_t bar) {
    if (!mn_major) {
        tma_load_2d(dst, map, k0, row0, bar);
    } else {
#pragma unroll
        for (int j = 0; j < ROWS / 32; ++j) tma_load_2d(dst + j * 4096, map, row0 + 32 * j, k0, bar);
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
          const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
    constexpr int TILE_B_BYTES = BN * BK * 4;
    constexpr int STAGE_BYTES = 2 * TILE_A_BYTES + 2 * TILE_B_BYTES;
    constexpr int HALF = BN / 2;                                               // columns owned by one epilogue warp
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // 128B swizzle atoms need 1 KB alignment
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE_BYTES);
    // bars[0..STAGES) full, [STAGES..2*STAGES) empty, then accumulator full [2], accumulator drained [2]; then the TMEM base
    const uint32_t bar_full = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_accfull = bar_empty + 8 * STAGES;
    const uint32_t bar_accfree = bar_accfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb_total = (p.K + BK - 1) / BK;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(nkb_total, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;                                         // >= 1 by construction of the grid
    const int nchunks = (nkb + KB_PER_CHUNK - 1) / KB_PER_CHUNK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmAh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBh) : "memory");
        if (p.three_pass) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmAl) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBl) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accfull + 8 * b, 1);                 // one tcgen05.commit
            mbar_init(bar_accfree + 8 * b, 8);                 // lane 0 of each of the 8 epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // TMEM: TWO accumulators of BN fp32 columns x 128 lanes (chunk ping-pong); this warp also frees them at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const uint32_t stage_tx = p.three_pass ? STAGE_BYTES : (TILE_A_BYTES + TILE_B_BYTES);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, stage_tx);
                const uint32_t st = smem_base + s * STAGE_BYTES;
                const int k0 = (kb_begin + i) * BK;
                load_operand<BM>(st, &tmAh, p.a_mn, m0, k0, full);
                load_operand<BN>(st + 2 * TILE_A_BYTES, &tmBh, p.b_mn, n0, k0, full);
                if (p.three_pass) {
                    load_operand<BM>(st + TILE_A_BYTES, &tmAl, p.a_mn, m0, k0, full);
                    load_operand<BN>(st + 2 * TILE_A_BYTES + TILE_B_BYTES, &tmBl, p.b_mn, n0, k0, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A/B TF32 [7,10)=[10,13)=2,
            // A/B major bits 15/16 (1 = MN-major), N>>3 in [17,23), M>>4 in [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            // K-major: 8-row groups 1 KB apart (SBO), K advance of 8 floats = +32 B inside the swizzle row
            // MN-major: 32-wide MN blocks 4 KB apart (LBO), 4-k-row atoms 512 B apart (SBO), K advance of 8 rows = +1 KB
            const uint32_t a_lbo = p.a_mn ? 4096u : 0u, b_lbo = p.b_mn ? 4096u : 0u;
            const uint32_t a_sbo = p.a_mn ? 512u : 1024u, b_sbo = p.b_mn ? 512u : 1024u;
            const uint32_t a_lt = p.a_mn ? 1u : 2u, b_lt = p.b_mn ? 1u : 2u;
            const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                const int chunk = i / KB_PER_CHUNK, in_chunk = i - chunk * KB_PER_CHUNK;
                const int buf = chunk & 1;
                if (in_chunk == 0 && chunk >= 2) {
                    // the epilogue warps must have drained this accumulator's previous chunk (use (chunk >> 1) - 1)
                    mbar_wait(bar_accfree + 8 * buf, (uint32_t)(((chunk >> 1) - 1) & 1));
                    tc_fence_after();
                }
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t acc = tmem_acc + (uint32_t)(buf * BN);
                const uint32_t st = smem_base + s * STAGE_BYTES;
                const uint32_t a_hi = st, a_lo = st + TILE_A_BYTES;
                const uint32_t b_hi = st + 2 * TILE_A_BYTES, b_lo = b_hi + TILE_B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint32_t accumulate = (in_chunk | k) != 0 ? 1u : 0u;      // a chunk starts from zero
                    const uint64_t dah = make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
                    const uint64_t dbh = make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
                    if (p.three_pass) {
                        const uint64_t dal = make_desc(a_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
                        const uint64_t dbl = make_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
                        // small terms first, so they are not absorbed by an already large accumulator
                        umma_tf32(acc, dal, dbh, idesc, accumulate);
                        umma_tf32(acc, dah, dbl, idesc, 1u);
                        umma_tf32(acc, dah, dbh, idesc, 1u);
                    } else {
                        umma_tf32(acc, dah, dbh, idesc, accumulate);
                    }
                }
                umma_commit(bar_empty + 8 * s);          // frees the stage once these MMAs have read it
                if (in_chunk == KB_PER_CHUNK - 1 || i == nkb - 1) umma_commit(bar_accfull + 8 * buf);   // chunk complete
            }
        }
    } else {
        // ===== epilogue: warps 2..9; a warp may only touch the TMEM lanes [32 * (warp % 4), +32); the two warps of a lane
        // quarter split the BN columns.  Per chunk: TMEM -> registers, release the accumulator, add (round to nearest). =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        float racc[HALF];
#pragma unroll
        for (int j = 0; j < HALF; ++j) racc[j] = 0.f;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(bar_accfull + 8 * buf, (uint32_t)((chunk >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * HALF);
#pragma unroll
            for (int c = 0; c < HALF; c += 32) {
                uint32_t v[32];
                TC_LD32(taddr + (uint32_t)c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 32 == HALF) {
                    // everything of this accumulator is in registers: hand it back to the MMA issuer before the adds
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_accfree + 8 * buf) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) racc[c + j] += __uint_as_float(v[j]);
            }
        }
        const int row = m0 + q * 32 + lane;
        const int nb = n0 + half * HALF;
        float* crow = p.C + (long long)blockIdx.z * p.split_stride + (long long)row * p.ldc + nb;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.split_stride % 4 == 0);
        if (row < p.M) {
#pragma unroll
            for (int c = 0; c < HALF; c += 32) {
                if (vec_ok && nb + c + 32 <= p.N) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(crow + c + j) = make_float4(racc[c + j], racc[c + j + 1], racc[c + j + 2], racc[c + j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + c + j < p.N) crow[c + j] = racc[c + j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)(2 * BN)) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D fp32 matrix, `inner` contiguous elements per row, `outer` rows of pitch `ld` elements; box {32, box_outer}
int make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_outer,
             bool mn_major) {
    EncodeTiledFn enc = get_encode();
    CAE_REQUIRE(enc != nullptr, "tc_gemm: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CAE_REQUIRE(r == CUDA_SUCCESS, "tc_gemm: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r, inner,
                outer, ld);
    return CAE_OK;
}

template <int BN, int STAGES>
int launch(const CaeTcGemm* g, const CUtensorMap* maps, const TcParams& p, dim3 grid, cudaStream_t st) {
    constexpr int smem = STAGES * (2 * TILE_A_BYTES + 2 * BN * BK * 4) + 1024 + 256;
    ensure_smem_limit(k_tc_gemm<BN, STAGES>, smem);
    k_tc_gemm<BN, STAGES><<<grid, NUM_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    return cae_check_launch("cae_tc_gemm");
}

}  // namespace

extern "C" int cae_tc_gemm(const CaeTcGemm* g, void* stream) {
    CAE_REQUIRE(g && g->a_hi && g->b_hi && g->C, "tc_gemm: null argument");
    CAE_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "tc_gemm: empty problem %dx%dx%d", g->M, g->N, g->K);
    CAE_REQUIRE((g->a_lo == nullptr) == (g->b_lo == nullptr), "tc_gemm: give both lo operands (3xTF32) or neither (1xTF32)");
    CAE_REQUIRE(g->lda % 4 == 0 && g->ldb % 4 == 0, "tc_gemm: operand pitches must be multiples of 4 floats (TMA: 16 bytes)");
    const float* ptrs[4] = {g->a_hi, g->a_lo, g->b_hi, g->b_lo};
    for (int i = 0; i < 4; ++i)
        CAE_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "tc_gemm: operand %d not 16-byte aligned", i);
    const int bn = (g->tile_n == 256) ? 256 : 128;
    const int splits = g->splits > 1 ? g->splits : 1;
    const int nkb = (g->K + BK - 1) / BK;
    TcParams p{};
    p.M = g->M; p.N = g->N; p.K = g->K;
    p.a_mn = g->a_mn_major ? 1 : 0;
    p.b_mn = g->b_mn_major ? 1 : 0;
    p.three_pass = g->a_lo != nullptr;
    p.kb_per_split = (nkb + splits - 1) / splits;
    p.C = g->C; p.ldc = g->ldc; p.split_stride = g->split_stride;
    const int zs = (nkb + p.kb_per_split - 1) / p.kb_per_split;      // every z slice owns at least one K block
    CAE_REQUIRE(zs == splits, "tc_gemm: %d splits leave empty K slices for K=%d (use at most %d)", splits, g->K, zs);
    CUtensorMap maps[4];
    int rc;
    const float* alo = g->a_lo ? g->a_lo : g->a_hi;
    const float* blo = g->b_lo ? g->b_lo : g->b_hi;
    // K-major: inner = K, rows = M|N, box {32, tile rows}; MN-major: inner = M|N, rows = K, box {32, 32}
    if (p.a_mn) {
        if ((rc = make_map(&maps[0], g->a_hi, g->M, g->K, g->lda, 32, true))) return rc;
        if ((rc = make_map(&maps[1], alo, g->M, g->K, g->lda, 32, true))) return rc;
    } else {
        if ((rc = make_map(&maps[0], g->a_hi, g->K, g->M, g->lda, BM, false))) return rc;
        if ((rc = make_map(&maps[1], alo, g->K, g->M, g->lda, BM, false))) return rc;
    }
    if (p.b_mn) {
        if ((rc = make_map(&maps[2], g->b_hi, g->N, g->K, g->ldb, 32, true))) return rc;
        if ((rc = make_map(&maps[3], blo, g->N, g->K, g->ldb, 32, true))) return rc;
    } else {
        if ((rc = make_map(&maps[2], g->b_hi, g->K, g->N, g->ldb, bn, false))) return rc;
        if ((rc = make_map(&maps[3], blo, g->K, g->N, g->ldb, bn, false))) return rc;
    }
    dim3 grid((g->N + bn - 1) / bn, (g->M + BM - 1) / BM, splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (bn == 256) return launch<256, 2>(g, maps, p, grid, st);
    return launch<128, 3>(g, maps, p, grid, st);
}

// ---- operand split: hi = x rounded to TF32, lo = x - hi rounded to TF32 (common.cuh: tf32_split) -------------------------
__global__ void k_tc_split(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = x[i];
        float h, l;
        tf32_split(v, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

extern "C" int cae_tc_split(const float* x, float* hi, float* lo, long long n, void* stream) {
    CAE_REQUIRE(x && hi && lo && n > 0, "tc_split: bad argument");
    const int grid = (int)min((long long)CAE_NUM_SMS * 8, (n + 255) / 256);
    k_tc_split<<<grid, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
    return cae_check_launch("cae_tc_split");
}
