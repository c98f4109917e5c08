// UP "tile pipeline" - transposed convolution (stride 2, pad 0, K = 3 | 4) for the wide thin layers, i.e. the forward pass of
// BASELINE configs[3]'s last three ConvTranspose2d layers (32->16 @127->255, 16->8 @255->511, 8->4 k4 @511->1024 with the
// fused sigmoid + MSE epilogue).  Same idea as k_down_tile: k_up3's operand rows arrive through a two-stage cp.async
// pipeline instead of loads at the top of every channel iteration.
//   CTA tile = 8 cell rows x 128 cells (thread = one cell row, 4 consecutive cells = 2 x 8 output pixels, COT channels);
//   stage    = the RAW rows of UT_CC input channels that the tile needs: 9 rows (one halo row above) x 132 columns (one
//              aligned chunk of halo to the left), zero-filled outside the plane;
//   the on-load affine (+ReLU) and the bounds mask (a transposed convolution's halo cells DO contribute: they must be
//   zero, not relu(k2)) are applied when a value moves from shared memory to registers.
// Epilogues and the reduction tail are those of k_up3; the target rows of the sigmoid + MSE epilogue are requested into
// L1 before the channel loop.
#pragma once
#include "conv_direct.cuh"
#include "conv_down_tile.cuh"

#define UT_ROWS 8
#define UT_STRIPS 32
#define UT_IW 132            // staged columns per row: 4 (aligned halo chunk) + 128
#define UT_IR (UT_ROWS + 1)
#define UT_CC 8              // input channels per stage

struct UpTilePlan {
    int tiles_y, tiles_x, ntiles;
    int nchunks;             // ceil(Cin / UT_CC)
};

template <int K, int COT>
__global__ void __launch_bounds__(CAE_NT, 2) k_up_tile(const ConvArgs a, const UpTilePlan p) {
    constexpr int KK = K * K;
    constexpr int STAGE = UT_CC * UT_IR * UT_IW;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                   // [ci][tap][COT]
    float* s_coef = s_w + a.Cin * KK * COT;              // [ci][4]: k0 - k2 -
    float* s_st = s_coef + ((a.Cin * 4 + 3) & ~3);       // two stages
    __shared__ EpiCh s_ech[COT];
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    const int co0 = blockIdx.y * COT;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    const int Hin = iv.H, Win = iv.W, Hout = a.out.H, Wout = a.out.W;

    for (int i = tid; i < a.Cin * KK * COT; i += CAE_NT) {
        const int j = i % COT, t = (i / COT) % KK, ci = i / (COT * KK);
        const int co = co0 + j;
        s_w[i] = co < a.Cout ? __ldg(a.w + ((size_t)ci * a.Cout + co) * KK + t) : 0.f;
    }
    for (int c = tid; c < a.Cin; c += CAE_NT) {
        const ChanCoef k = load_coef(a.in, c);
        s_coef[4 * c] = k.k0; s_coef[4 * c + 2] = k.k2;
    }
    if (tid < COT) s_ech[tid] = epi_load_channel(a.epi, min(co0 + tid, a.Cout - 1), co0 + tid < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;

    const int per_sample = p.tiles_y * p.tiles_x;
    constexpr int chunks_row = UT_IW / 4;
    auto fetch = [&](int tile, int chunk, int buf) {
        const int n = tile / per_sample, tr = tile - n * per_sample;
        const int tyi = tr / p.tiles_x, txi = tr - tyi * p.tiles_x;
        const int r0 = tyi * UT_ROWS - 1, c0 = txi * UT_STRIPS * 4 - 4;
        float* dst0 = s_st + buf * STAGE;
        const int ci0 = chunk * UT_CC, cn = min(UT_CC, a.Cin - ci0);
        const float* base = iv.p + in_base + (long long)n * iv.sN;
        for (int q = tid; q < cn * UT_IR * chunks_row; q += CAE_NT) {
            const int row = q / chunks_row, xq = q - row * chunks_row;
            const int cl = row / UT_IR, rr = row - cl * UT_IR;
            const int gr = r0 + rr, gc = c0 + 4 * xq;
            const bool ok = gr >= 0 && gr < Hin && gc >= 0 && gc < iv.ld;
            dt_cp16(dst0 + row * UT_IW + 4 * xq, ok ? base + (long long)(ci0 + cl) * iv.sC + (long long)gr * iv.ld + gc : base, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int buf = 0;
    if ((int)blockIdx.x < p.ntiles) fetch(blockIdx.x, 0, 0);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int n = tile / per_sample, tr = tile - n * per_sample;
        const int tyi = tr / p.tiles_x, txi = tr - tyi * p.tiles_x;
        const int qy = tyi * UT_ROWS + ty, qx0 = (txi * UT_STRIPS + tx) * 4;
        const bool live = 2 * qy < Hout && 2 * qx0 < Wout;
        if (a.epi.mode == CAE_EPI_SIGMOID_MSE && live) {
            const CaeView& t = a.epi.target.t0;
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < a.Cout)
#pragma unroll
                    for (int py = 0; py < 2; ++py)
                        if (2 * qy + py < Hout)
                            pf_l1(t.p + tgt_base + (long long)n * t.sN + (long long)(co0 + j) * t.sC + (long long)(2 * qy + py) * t.ld + 2 * qx0);
        }
        float acc[COT][2][8];
#pragma unroll
        for (int j = 0; j < COT; ++j)
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[j][py][e] = 0.f;
        // which of the five columns qx0 - 1 .. qx0 + 3 and of the two rows qy, qy - 1 exist
        bool cok[5], rok[2];
#pragma unroll
        for (int i = 0; i < 5; ++i) cok[i] = qx0 - 1 + i >= 0 && qx0 - 1 + i < Win;
        rok[0] = qy < Hin;
        rok[1] = qy - 1 >= 0 && qy - 1 < Hin;
        for (int chunk = 0; chunk < p.nchunks; ++chunk) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                 // stage `buf` landed; the other stage is free (first time: s_w / s_coef / s_ech visible)
            if (chunk + 1 < p.nchunks) fetch(tile, chunk + 1, buf ^ 1);
            else if (tile + (int)gridDim.x < p.ntiles) fetch(tile + gridDim.x, 0, buf ^ 1);
            const int ci0 = chunk * UT_CC, cn = min(UT_CC, a.Cin - ci0);
            for (int cl = 0; cl < cn; ++cl) {
                const int ci = ci0 + cl;
                const float k0 = s_coef[4 * ci], k2 = s_coef[4 * ci + 2];
                // stage row 0 holds input row tile_row0 - 1: thread row ty reads rows ty + 1 (jy = 0) and ty (jy = 1)
                const float* st = s_st + buf * STAGE + (cl * UT_IR + ty + 1) * UT_IW + 4 * tx + 3;
                float v[2][5];
#pragma unroll
                for (int jy = 0; jy < 2; ++jy) {
                    const float* rp = st - jy * UT_IW;
                    const float4 q = *reinterpret_cast<const float4*>(rp + 1);
                    const float raw[5] = {rp[0], q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        float t = fmaf(raw[i], k0, k2);
                        if (a.in.relu) t = fmaxf(t, 0.f);
                        v[jy][i] = (rok[jy] && cok[i]) ? t : 0.f;
                    }
                }
                const float* wp = s_w + ci * KK * COT;
#pragma unroll
                for (int py = 0; py < 2; ++py)
#pragma unroll
                    for (int jy = 0; jy < 2; ++jy) {
                        const int ky = py + 2 * jy;
                        if (ky < K) {
#pragma unroll
                            for (int px = 0; px < 2; ++px)
#pragma unroll
                                for (int jx = 0; jx < 2; ++jx) {
                                    const int kx = px + 2 * jx;
                                    if (kx < K) {
                                        float wv[COT];
#pragma unroll
                                        for (int j = 0; j < COT; ++j) wv[j] = wp[(ky * K + kx) * COT + j];
#pragma unroll
                                        for (int cx = 0; cx < 4; ++cx) {
                                            const float xv = v[jy][cx + 1 - jx];
#pragma unroll
                                            for (int j = 0; j < COT; ++j)
                                                acc[j][py][2 * cx + px] = fmaf(xv, wv[j], acc[j][py][2 * cx + px]);
                                        }
                                    }
                                }
                        }
                    }
            }
            buf ^= 1;
        }
        if (live) {
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const int oy = 2 * qy + py;
                if (oy < Hout) {
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < a.Cout) {
                            const EpiCh ech = s_ech[j];
                            epi_strip<8>(a.epi, a.out, ech, n, co0 + j, oy, 2 * qx0, acc[j][py], tgt_base, a.inv_count, s1[j], s2[j]);
                        }
                }
            }
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}
