// DOWN "tile pipeline" - strided convolution (stride 2, pad 0, K = 3 | 4) for the wide thin layers, i.e. the input gradient of
// BASELINE configs[3]'s last three ConvTranspose2d layers (16->32 @255->127, 8->16 @511->255, 4->8 k4 @1024->511).
// k_down3 loads its operand rows straight from global memory at the top of every channel iteration: with 16 warps per SM
// (128 registers) nothing hides those misses (ncu: long_scoreboard 4.5 - 5.7 warps per issue, issue slots 39 - 46 % busy,
// DRAM 25 %).  Here the rows come through a two-stage cp.async pipeline:
//   CTA tile = 8 output rows x 128 output columns (thread = one row, 4 consecutive pixels, COT output channels);
//   stage    = the RAW rows of ONE input channel (t0 and, where present, t1) that the tile needs: 2*8 + K - 2 rows x 260
//              columns, zero-filled outside the plane; stage s+1 is in flight while stage s is accumulated;
//   the on-load affine (+ReLU) is applied when a value moves from shared memory to registers.
// Outputs outside the plane are never stored, and valid outputs never touch input outside the plane (pad 0), so the
// zero fill needs no mask.  Epilogues and the reduction tail are those of k_down3 (same partial-row layout).
#pragma once
#include "conv_direct.cuh"

#define DT_ROWS 8
#define DT_STRIPS 32
#define DT_IW 260            // staged columns per row: 2*128 + K - 2 rounded up to whole 16-byte chunks

struct DownTilePlan {
    int tiles_y, tiles_x, ntiles;
    int IR;                  // staged rows per channel: 2*DT_ROWS + K - 2
    int stage_fl;            // floats of one stage: ntens * IR * DT_IW
    int ntens;               // 1 | 2
};

__device__ __forceinline__ void dt_cp16(float* dst, const float* src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int K, int COT>
__global__ void __launch_bounds__(CAE_NT, 2) k_down_tile(const ConvArgs a, const DownTilePlan p) {
    constexpr int KK = K * K;
    constexpr int NVV = 8 + K - 2;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                   // [ci][tap][COT]
    float* s_coef = s_w + a.Cin * KK * COT;              // [ci][4]: k0 k1 k2 -
    float* s_st = s_coef + ((a.Cin * 4 + 3) & ~3);       // two stages
    __shared__ EpiCh s_ech[COT];
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    const int co0 = blockIdx.y * COT;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const int Hin = iv.H, OH = a.out.H, OW = a.out.W;
    const bool has_t1 = a.in.t1 != nullptr;

    for (int i = tid; i < a.Cin * KK * COT; i += CAE_NT) {
        const int j = i % COT, t = (i / COT) % KK, ci = i / (COT * KK);
        const int co = co0 + j;
        s_w[i] = co < a.Cout ? __ldg(a.w + ((size_t)co * a.Cin + ci) * KK + t) : 0.f;
    }
    for (int c = tid; c < a.Cin; c += CAE_NT) {
        const ChanCoef k = load_coef(a.in, c);
        s_coef[4 * c] = k.k0; s_coef[4 * c + 1] = k.k1; s_coef[4 * c + 2] = k.k2;
    }
    if (tid < COT) s_ech[tid] = epi_load_channel(a.epi, min(co0 + tid, a.Cout - 1), co0 + tid < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;

    const int per_sample = p.tiles_y * p.tiles_x;
    const int chunks_row = DT_IW / 4;
    auto fetch = [&](int tile, int ci, int buf) {
        const int n = tile / per_sample, tr = tile - n * per_sample;
        const int tyi = tr / p.tiles_x, txi = tr - tyi * p.tiles_x;
        const int r0 = 2 * tyi * DT_ROWS, c0 = 2 * txi * DT_STRIPS * 4;
        float* dst0 = s_st + buf * p.stage_fl;
        for (int t = 0; t < p.ntens; ++t) {
            const float* base = (t ? a.in.t1 : iv.p) + in_base + (long long)n * iv.sN + (long long)ci * iv.sC;
            float* dst = dst0 + t * p.IR * DT_IW;
            for (int q = tid; q < p.IR * chunks_row; q += CAE_NT) {
                const int rr = q / chunks_row, xq = q - rr * chunks_row;
                const int gr = r0 + rr, gc = c0 + 4 * xq;
                const bool ok = gr < Hin && gc < iv.ld;
                dt_cp16(dst + rr * DT_IW + 4 * xq, ok ? base + (long long)gr * iv.ld + gc : base, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // stages in order: (tile of this CTA, ci)
    int buf = 0;
    if ((int)blockIdx.x < p.ntiles) fetch(blockIdx.x, 0, 0);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int n = tile / per_sample, tr = tile - n * per_sample;
        const int tyi = tr / p.tiles_x, txi = tr - tyi * p.tiles_x;
        const int oy = tyi * DT_ROWS + ty, ox0 = (txi * DT_STRIPS + tx) * 4;
        unsigned long long acc2[COT / 2][4];             // channel pairs (j, j + 1): one FFMA2 per pair (wgt_fma2, bit-identical to fmaf)
#pragma unroll
        for (int j = 0; j < COT / 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2[j][e] = 0ull;
        for (int ci = 0; ci < a.Cin; ++ci) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                 // stage `buf` landed; everyone finished reading the other stage (and, the first time, s_w / s_coef are visible)
            if (ci + 1 < a.Cin) fetch(tile, ci + 1, buf ^ 1);
            else if (tile + (int)gridDim.x < p.ntiles) fetch(tile + gridDim.x, 0, buf ^ 1);
            const float k0 = s_coef[4 * ci], k1 = s_coef[4 * ci + 1], k2 = s_coef[4 * ci + 2];
            const float* st = s_st + buf * p.stage_fl + (2 * ty) * DT_IW + 8 * tx;
            const float* wp = s_w + ci * KK * COT;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const float* rp = st + ky * DT_IW;
                float v[NVV];
                {
                    const float4 q0 = *reinterpret_cast<const float4*>(rp), q1 = *reinterpret_cast<const float4*>(rp + 4);
                    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
#pragma unroll
                    for (int i = 8; i < NVV; ++i) v[i] = rp[i];
                }
#pragma unroll
                for (int i = 0; i < NVV; ++i) v[i] = fmaf(v[i], k0, k2);
                if (has_t1) {
                    const float* rq = rp + p.IR * DT_IW;
                    const float4 q0 = *reinterpret_cast<const float4*>(rq), q1 = *reinterpret_cast<const float4*>(rq + 4);
                    v[0] = fmaf(q0.x, k1, v[0]); v[1] = fmaf(q0.y, k1, v[1]); v[2] = fmaf(q0.z, k1, v[2]); v[3] = fmaf(q0.w, k1, v[3]);
                    v[4] = fmaf(q1.x, k1, v[4]); v[5] = fmaf(q1.y, k1, v[5]); v[6] = fmaf(q1.z, k1, v[6]); v[7] = fmaf(q1.w, k1, v[7]);
#pragma unroll
                    for (int i = 8; i < NVV; ++i) v[i] = fmaf(rq[i], k1, v[i]);
                }
                if (a.in.relu) {
#pragma unroll
                    for (int i = 0; i < NVV; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                unsigned long long v2[NVV];
#pragma unroll
                for (int i = 0; i < NVV; ++i) v2[i] = wgt_pk(v[i], v[i]);
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    unsigned long long wv[COT / 2];
#pragma unroll
                    for (int j = 0; j < COT; j += 4) {
                        const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wp + (ky * K + kx) * COT + j);
                        wv[j / 2] = w4.x; wv[j / 2 + 1] = w4.y;
                    }
#pragma unroll
                    for (int cx = 0; cx < 4; ++cx)
#pragma unroll
                        for (int j = 0; j < COT / 2; ++j) acc2[j][cx] = wgt_fma2(v2[2 * cx + kx], wv[j], acc2[j][cx]);
                }
            }
            buf ^= 1;
        }
        float acc[COT][4];
#pragma unroll
        for (int j = 0; j < COT / 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) wgt_upk(acc2[j][e], acc[2 * j][e], acc[2 * j + 1][e]);
        if (oy < OH && ox0 < OW) {
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < a.Cout) {
                    const EpiCh ech = s_ech[j];
                    epi_strip<4>(a.epi, a.out, ech, n, co0 + j, oy, ox0, acc[j], 0ll, a.inv_count, s1[j], s2[j]);
                }
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}
