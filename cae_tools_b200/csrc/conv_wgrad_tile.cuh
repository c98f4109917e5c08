// WGRAD "tile resident" - weight gradient of the wide, thin stride-2 layers (BASELINE configs[3]: 32->16 @127->255,
// 16->8 @255->511, 8->4 k4 @511->1024).
//   G[cs][cb][ky][kx] = sum_{n,i,j} S(n,cs,i,j) * B(n,cb,2i+ky,2j+kx)          (pad 0, stride 2, K = 3 | 4)
// The direct kernel (k_wgrad3) gives every (cs tile, cb tile) pair its own CTAs, so both operands are fetched once per
// pair (8 ... 64 times for these layers).  Here a CTA stages one spatial tile of ALL channels of both operands in shared
// memory - channel innermost, so that a thread fetches its CST small values with one 16/8-byte load and a CBT = 2 pair
// of big values with one 8-byte load - and every one of the Cs*Cb*K*K weight elements is accumulated from that copy.
// Thread roles: 256 = n_cbt x PSH x n_cst x 4;  (cst, cbt) = the register tile (CST x CBT x K*K accumulators),
// (ps_hi, ps_lo) = which of the tile's positions the thread walks (PS = 4*PSH position lanes per register tile).
// Accumulators live across the CTA's tiles; one partial row per CTA, summed in row order by the last CTA (deterministic).
#pragma once
#include "conv_family.cuh"

#define WGT_TW 32
#define WGT_BCR 72            // columns of a raw big-tile row: 2*32 + K - 2 rounded up to whole 16-byte chunks

struct WgTilePlan {
    int TH;                 // tile rows (small positions); the tile is TH x 32
    int tiles_y, tiles_x;   // tiles per sample
    int ntiles;             // N * tiles_y * tiles_x
    int CsP, CbP;           // channel pitches of the two cooked tiles (Cs + 4, Cb + 2: conflict-free for the access patterns)
    int BR, BC;             // rows / columns of the big tile: 2*TH + K - 2, 2*32 + K - 2
    int n_cst, n_cbt;       // register tiles along the two channel axes
    int PSH;                // PS / 4
    int TH_shift;           // log2(TH)
    int inv_BR;             // ceil(65536 / BR): v / BR == (v * inv_BR) >> 16 for the row counts that occur
    int SCH, BCH;           // channel strides of the raw tiles (== 8 mod 32: the cook pass reads 4 channels x 8 columns per warp)
    int raw_s, raw_b;       // floats of one raw small / big tensor tile (Cs*SCH, Cb*BCH)
    int cooked;             // float offset of the cooked tiles behind the raw ones
};

__device__ __forceinline__ void wgt_cp16(float* dst, const float* src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 16 : 0;                       // 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// Pipeline per CTA tile: (1) cp.async the RAW rows of both operands (t0 and, where present, t1) into planar shared tiles -
// issued for tile i+1 before tile i is accumulated, so the HBM latency hides behind the FMAs; (2) "cook" raw -> cooked:
// on-load affine (+ReLU), bounds mask of the small operand, transpose to channel-innermost; (3) accumulate from cooked.
template <int K, int CST, int CBT>
__global__ void __launch_bounds__(CAE_NT, 2) k_wgrad_tile(const WgradArgs a, const WgTilePlan p) {
    static_assert(CBT == 2 && (CST == 2 || CST == 4), "register tile");
    constexpr int KK = K * K;
    constexpr int NACC = CST * CBT * KK;
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_coef[3 * 96];                     // k0 k1 k2 of the Cs small and Cb big channels
    const int NPOS = p.TH * WGT_TW;
    const bool s_t1 = a.sm.t1 != nullptr, b_t1 = a.bg.t1 != nullptr;
    float* r_s = smem;                                   // raw small: [1 | 2][Cs][SCH]
    float* r_b = r_s + (s_t1 ? 2 : 1) * p.raw_s;         // raw big:   [1 | 2][Cb][BCH]
    float* s_s = smem + p.cooked;                        // cooked small [NPOS][CsP]
    float* s_b = s_s + NPOS * p.CsP;                     // cooked big   [BR][BC][CbP]
    const int tid = threadIdx.x;
    const int ps_lo = tid & 3;
    const int cst = (tid >> 2) % p.n_cst;
    const int r_ = (tid >> 2) / p.n_cst;
    const int ps_hi = r_ % p.PSH, cbt = r_ / p.PSH;
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    const int Hs = sv.H, Ws = sv.W, Hb = bv.H;
    const int Cs = a.Cs, Cb = a.Cb;

    for (int c = tid; c < Cs + Cb; c += CAE_NT) {
        const ChanCoef k = c < Cs ? load_coef(a.sm, c) : load_coef(a.bg, c - Cs);
        s_coef[3 * c] = k.k0; s_coef[3 * c + 1] = k.k1; s_coef[3 * c + 2] = k.k2;
    }

    float acc[CST][CBT][KK];
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) acc[x][y][t] = 0.f;

    const int per_sample = p.tiles_y * p.tiles_x;
    const int lane = tid & 31, warp = tid >> 5, c_lo = lane >> 3, x_lo = lane & 7;
    const int niter = NPOS / (4 * p.PSH);
    const float* sp0 = s_s + cst * CST;
    const float* bp0 = s_b + cbt * CBT;

    auto fetch = [&](int tile) {
        const int n = tile / per_sample, tr = tile - n * per_sample;
        const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
        const int i0 = ty * p.TH, j0 = tx * WGT_TW;
        for (int t = 0; t < (s_t1 ? 2 : 1); ++t) {
            const float* base = (t ? a.sm.t1 : sv.p) + sbase + (long long)n * sv.sN;
            float* dst = r_s + t * p.raw_s;
            for (int q = tid; q < Cs * p.TH * 8; q += CAE_NT) {
                const int xq = q & 7, y = (q >> 3) & (p.TH - 1), c = q >> (p.TH_shift + 3);
                const int gi = i0 + y, gj = j0 + 4 * xq;
                const bool ok = gi < Hs && gj < sv.ld;
                wgt_cp16(dst + c * p.SCH + y * WGT_TW + 4 * xq,
                         ok ? base + (long long)c * sv.sC + (long long)gi * sv.ld + gj : base, ok);
            }
        }
        for (int t = 0; t < (b_t1 ? 2 : 1); ++t) {
            const float* base = (t ? a.bg.t1 : bv.p) + bbase + (long long)n * bv.sN;
            float* dst = r_b + t * p.raw_b;
            for (int q = tid; q < Cb * p.BR * (WGT_BCR / 4); q += CAE_NT) {
                const int row = q / (WGT_BCR / 4), xq = q - row * (WGT_BCR / 4);
                const int c = (row * p.inv_BR) >> 16, r = row - c * p.BR;
                const int gr = 2 * i0 + r, gc = 2 * j0 + 4 * xq;
                const bool ok = gr < Hb && gc < bv.ld;
                wgt_cp16(dst + c * p.BCH + r * WGT_BCR + 4 * xq,
                         ok ? base + (long long)c * bv.sC + (long long)gr * bv.ld + gc : base, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if ((int)blockIdx.x < p.ntiles) fetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int n = tile / per_sample, tr = tile - n * per_sample;
        const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
        const int i0 = ty * p.TH, j0 = tx * WGT_TW;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                 // raw tiles landed; everyone is done accumulating the previous tile
        // ---- cook the small tile.  A warp task = 4 channels x one tile row; a lane owns (channel c_lo, columns x_lo + 8*xo):
        // conflict-free on both sides (raw channel stride == 8 mod 32; cooked channel pitch Cs + 4).
        for (int task = warp; task < (Cs >> 2) * p.TH; task += CAE_NWARP) {
            const int cg = task >> p.TH_shift, y = task & (p.TH - 1);
            const int c = cg * 4 + c_lo;
            const float k0 = s_coef[3 * c], k1 = s_coef[3 * c + 1], k2 = s_coef[3 * c + 2];
            const float* src = r_s + c * p.SCH + y * WGT_TW;
            float* dst = s_s + (y * WGT_TW) * p.CsP + c;
            const bool rok = i0 + y < Hs;
#pragma unroll
            for (int xo = 0; xo < 4; ++xo) {
                const int x = xo * 8 + x_lo;
                float v = fmaf(src[x], k0, k2);
                if (s_t1) v = fmaf(src[p.raw_s + x], k1, v);
                if (a.sm.relu) v = fmaxf(v, 0.f);
                dst[x * p.CsP] = (rok && j0 + x < Ws) ? v : 0.f;
            }
        }
        // ---- cook the big tile: task = 4 channels x one tile row x three of its nine column octets.  No bounds mask:
        // out-of-range big elements only ever meet out-of-range (zeroed) small positions, and they were zero-filled.
        for (int task = warp; task < (Cb >> 2) * p.BR * 3; task += CAE_NWARP) {
            const int rest = task / 3, xg = task - rest * 3;
            const int cg = (rest * p.inv_BR) >> 16, r = rest - cg * p.BR;
            const int c = cg * 4 + c_lo;
            const float k0 = s_coef[3 * (Cs + c)], k1 = s_coef[3 * (Cs + c) + 1], k2 = s_coef[3 * (Cs + c) + 2];
            const float* src = r_b + c * p.BCH + r * WGT_BCR;
            float* dst = s_b + (r * p.BC) * p.CbP + c;
#pragma unroll
            for (int xo = 0; xo < 3; ++xo) {
                const int x = (xg * 3 + xo) * 8 + x_lo;
                if (x < p.BC) {
                    float v = fmaf(src[x], k0, k2);
                    if (b_t1) v = fmaf(src[p.raw_b + x], k1, v);
                    if (a.bg.relu) v = fmaxf(v, 0.f);
                    dst[x * p.CbP] = v;
                }
            }
        }
        __syncthreads();                                 // cooked tiles ready; raw tiles free again
        if (tile + (int)gridDim.x < p.ntiles) fetch(tile + gridDim.x);
        // ---- accumulate
        for (int it = 0; it < niter; ++it) {
            const int pos = ((it * p.PSH + ps_hi) << 2) + ps_lo;
            const int y = pos >> 5, x = pos & 31;
            float sval[CST];
            if constexpr (CST == 4) {
                const float4 t = *reinterpret_cast<const float4*>(sp0 + pos * p.CsP);
                sval[0] = t.x; sval[1] = t.y; sval[2] = t.z; sval[3] = t.w;
            } else {
                const float2 t = *reinterpret_cast<const float2*>(sp0 + pos * p.CsP);
                sval[0] = t.x; sval[1] = t.y;
            }
            const float* bp = bp0 + ((2 * y) * p.BC + 2 * x) * p.CbP;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float2 b = *reinterpret_cast<const float2*>(bp + (ky * p.BC + kx) * p.CbP);
#pragma unroll
                    for (int c = 0; c < CST; ++c) {
                        acc[c][0][ky * K + kx] = fmaf(sval[c], b.x, acc[c][0][ky * K + kx]);
                        acc[c][1][ky * K + kx] = fmaf(sval[c], b.y, acc[c][1][ky * K + kx]);
                    }
                }
            }
        }
    }
    // ---- reduce over the position lanes: ps_lo by shuffle, ps_hi through shared memory (fixed order)
    __syncthreads();
    float* red = smem;                                   // [64][NACC] (the tiles are no longer needed)
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                float v = acc[x][y][t];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (ps_lo == 0) red[(tid >> 2) * NACC + (x * CBT + y) * KK + t] = v;
            }
    __syncthreads();
    const int nelem = Cs * Cb * KK;
    for (int o = tid; o < nelem; o += CAE_NT) {
        const int pt = o / NACC, idx = o - pt * NACC;
        const int ob = pt / p.n_cst, os = pt - ob * p.n_cst;
        float s = 0.f;
        for (int h = 0; h < p.PSH; ++h) s += red[((ob * p.PSH + h) * p.n_cst + os) * NACC + idx];
        const int cs = os * CST + idx / (CBT * KK), cb = ob * CBT + (idx / KK) % CBT, t = idx % KK;
        a.partials[(size_t)blockIdx.x * nelem + ((size_t)cs * Cb + cb) * KK + t] = s;
    }
    if (cae_last_block(a.ticket)) {
        const int rows = gridDim.x;
        for (int e = tid; e < nelem; e += CAE_NT) {
            float s = 0.f;
            int r = 0;
            for (; r + 3 < rows; r += 4) {
                float v0 = __ldcg(a.partials + (size_t)r * nelem + e), v1 = __ldcg(a.partials + (size_t)(r + 1) * nelem + e);
                float v2 = __ldcg(a.partials + (size_t)(r + 2) * nelem + e), v3 = __ldcg(a.partials + (size_t)(r + 3) * nelem + e);
                s += v0; s += v1; s += v2; s += v3;
            }
            for (; r < rows; ++r) s += __ldcg(a.partials + (size_t)r * nelem + e);
            a.grad[e] = s;
        }
    }
}
