// Host-side helpers shared by the translation units of libcae_b200.so (argument checks, kernel-selection mask).
#pragma once
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_set>
#include "common.cuh"

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

static int check_view(const CaeView& v, const char* name) {
    CAE_REQUIRE(v.p != nullptr, "%s: null pointer", name);
    CAE_REQUIRE(v.N > 0 && v.C > 0 && v.H > 0 && v.W > 0, "%s: empty tensor %dx%dx%dx%d", name, v.N, v.C, v.H, v.W);
    CAE_REQUIRE(v.ld >= v.W, "%s: ld %d < W %d", name, v.ld, v.W);
    CAE_REQUIRE((long long)v.N * v.C * v.H * v.W < (1ll << 31), "%s: tensor too large for 32-bit positions", name);
    return CAE_OK;
}

static const int kTileSmemBudget = 72 * 1024;
static const int kTileSmemMax = 100 * 1024;
// kernel selection mask: bit 0 tiled up/down (k_up2/k_down2), bit 1 position-parallel wgrad (k_wgrad2a),
// bit 2 GEMM-like wgrad (k_wgrad2b).  cae_set_kernel_generation(1) = generic kernels only, (2) = default mask.
#define CAE_V2_UPDOWN 1
#define CAE_V2_WGRAD_A 2
#define CAE_V2_WGRAD_B 4
#define CAE_V2_UPDOWN_WIDE 8   // tiled up/down also for wide layers (register tile of 2/4 positions)
#define CAE_V3_DIRECT 16       // vectorised direct kernels for wide thin layers (k_up3 / k_down3)
// warp-per-element weight gradient for layers with few positions (k_wgrad_small).  OFF by default: alone it is 2x faster
// than the tiled kernels (33 -> 14 us), but its 1000+ CTAs take the SMs away from the input-gradient chain that runs
// beside it on the main stream - the whole step got slower (unet 341 -> 350 us, conv 492 -> 528 us).
#define CAE_WGRAD_SMALL 32
// tile-resident weight gradient of the wide thin layers (k_wgrad_tile): both operands are fetched once per CTA tile
// instead of once per channel-tile pair
#define CAE_WGRAD_TILE 64
// cp.async tile pipeline for the strided conv of the wide thin layers (k_down_tile) instead of the direct-load k_down3
#define CAE_DOWN_TILE 128
#define CAE_UP_TILE 256      // the same for the transposed conv (k_up_tile instead of k_up3)
extern int g_cae_mask;                 // defined in capi.cu
#define g_mask g_cae_mask
#define g_use_v2 (g_mask & CAE_V2_UPDOWN)

static inline int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static inline int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
static inline int roundup4(int v) { return (v + 3) & ~3; }

// opt in to > 48 KB dynamic shared memory, once per kernel (keyed by the function address)
template <typename K>
static void ensure_smem(K kernel) {
    static std::mutex mu;
    static std::unordered_set<const void*> seen;
    std::lock_guard<std::mutex> lock(mu);
    if (seen.insert((const void*)kernel).second)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemMax);
}

static bool aligned4(const void* p, int ld, long long sC, long long sN) {
    return p && ((uintptr_t)p % 16 == 0) && ld % 4 == 0 && sC % 4 == 0 && sN % 4 == 0;
}
static bool view_aligned(const CaeView& v) { return aligned4(v.p, v.ld, v.sC, v.sN); }
static bool src_aligned(const CaeSrc& s) {
    if (!view_aligned(s.t0)) return false;
    if (s.t1 && (uintptr_t)s.t1 % 16 != 0) return false;
    if (s.cursor && s.cursor_stride % 4 != 0) return false;
    return true;
}

// opt in to `bytes` of dynamic shared memory for one kernel (once per kernel)
template <typename K>
static void ensure_smem_limit(K kernel, int bytes) {
    static std::mutex mu;
    static std::unordered_set<const void*> seen;
    std::lock_guard<std::mutex> lock(mu);
    if (seen.insert((const void*)kernel).second)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
