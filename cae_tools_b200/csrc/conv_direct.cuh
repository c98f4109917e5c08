// v3 "direct" kernels for the wide, thin layers (few channels, long rows; stride 2, square 3x3 / 4x4 kernels,
// no padding): HBM-bound, so the aim is a minimal instruction stream per byte.
//   * a thread owns a strip of 4 consecutive cells / output pixels of one row for COT channels;
//   * operands come straight from global memory as aligned float4 (row pitches are multiples of 4 floats),
//     the one halo column as a scalar (an L1 hit: the neighbouring lane's vector holds it);
//   * weights are broadcast from shared memory; results leave as float4.
#pragma once
#include "conv_family.cuh"

struct StripPlan {
    int RP;          // rows per sample on the flattened row axis
    int NS;          // strips per row
    int units;       // N * RP * NS
};

// software prefetch into L1: these kernels hold 16 warps per SM (128 registers) and every channel iteration used to start with
// an L1 miss (ncu: long_scoreboard 2.2 - 5.7 warps per issue, issue slots 40 - 55 % busy).  The rows of the NEXT channel are
// requested while the current one is accumulated; no registers are held.
__device__ __forceinline__ void pf_l1(const float* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pf_src(const CaeSrc& s, long long off) {
    pf_l1(s.t0.p + off);
    if (s.t1) pf_l1(s.t1 + off);
}

__device__ __forceinline__ float xf1(float v, const ChanCoef& k, bool relu) {
    v = fmaf(v, k.k0, k.k2);
    return relu ? fmaxf(v, 0.f) : v;
}

// one channel's four consecutive values starting at element offset `off` (16-byte aligned), transformed
__device__ __forceinline__ void load4(const CaeSrc& s, long long off, const ChanCoef& k, float (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(s.t0.p + off));
    v[0] = fmaf(a.x, k.k0, k.k2); v[1] = fmaf(a.y, k.k0, k.k2); v[2] = fmaf(a.z, k.k0, k.k2); v[3] = fmaf(a.w, k.k0, k.k2);
    if (s.t1) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(s.t1 + off));
        v[0] = fmaf(b.x, k.k1, v[0]); v[1] = fmaf(b.y, k.k1, v[1]); v[2] = fmaf(b.z, k.k1, v[2]); v[3] = fmaf(b.w, k.k1, v[3]);
    }
    if (s.relu) {
        v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f);
    }
}

// epilogue for 8 (UP) or 4 (DOWN) consecutive outputs of one row starting at a 16-byte aligned column
template <int NE>
__device__ __forceinline__ void epi_strip(const CaeEpilogue& e, const CaeView& out, const EpiCh& ch, int n, int co, int oy,
                                          int ox0, float (&acc)[NE], long long tgt_base, float inv_count, float& s1,
                                          float& s2) {
    const int Wout = out.W;
    const long long ro = (long long)n * out.sN + (long long)co * out.sC + (long long)oy * out.ld + ox0;
    float* orow = out.p + ro;
    float res[NE];
    bool write = true;
    switch (e.mode) {
        case CAE_EPI_PLAIN:
#pragma unroll
            for (int i = 0; i < NE; ++i) res[i] = acc[i] + ch.bias;
            break;
        case CAE_EPI_STATS:
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                float v = acc[i] + ch.bias;
                res[i] = v;
                if (ox0 + i < Wout) {
                    s1 += v;
                    s2 = fmaf(v, v, s2);
                }
            }
            break;
        case CAE_EPI_MASKSTATS: {
            const CaeView& a = e.act;
            const float* arow = a.p + ((long long)n * a.sN + (long long)co * a.sC + (long long)oy * a.ld + ox0);
#pragma unroll
            for (int q = 0; q < NE / 4; ++q) {
                const float4 y4 = __ldg(reinterpret_cast<const float4*>(arow + 4 * q));
                const float yp[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float z = fmaf(yp[i], ch.scale, ch.shift);
                    float dz = z > 0.f ? acc[4 * q + i] : 0.f;
                    res[4 * q + i] = dz;
                    if (ox0 + 4 * q + i < Wout) {
                        s1 += dz;
                        s2 = fmaf(dz, (yp[i] - ch.mean) * ch.invstd, s2);
                    }
                }
            }
        } break;
        case CAE_EPI_SIGMOID:
#pragma unroll
            for (int i = 0; i < NE; ++i) res[i] = cae_fast_sigmoid(acc[i] + ch.bias);
            break;
        case CAE_EPI_SIGMOID_MSE: {
            const CaeView& t = e.target.t0;
            const float* trow = t.p + (tgt_base + (long long)n * t.sN + (long long)co * t.sC + (long long)oy * t.ld + ox0);
#pragma unroll
            for (int q = 0; q < NE / 4; ++q) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + 4 * q));
                const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float yh = cae_fast_sigmoid(acc[4 * q + i] + ch.bias);
                    float d = yh - fmaf(tv[i], ch.tk.k0, ch.tk.k2);
                    float dz = 2.f * d * inv_count * yh * (1.f - yh);
                    if (ox0 + 4 * q + i < Wout) {
                        s2 = fmaf(d, d, s2);
                        s1 += dz;
                    }
                    res[4 * q + i] = e.write_mode == 1 ? yh : dz;
                }
            }
            write = e.write_mode != 2;
        } break;
    }
    if (write) {
#pragma unroll
        for (int q = 0; q < NE / 4; ++q) {
            if (ox0 + 4 * q + 3 < Wout) {
                *reinterpret_cast<float4*>(orow + 4 * q) = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (ox0 + 4 * q + i < Wout) orow[4 * q + i] = res[4 * q + i];
            }
        }
    }
}

// =======================================================================================
// UP v3: transposed conv, stride 2, pad 0.  Thread = 4 cells (8 x 2 output pixels) x COT channels.
// =======================================================================================
template <int K, int COT>
__global__ void __launch_bounds__(CAE_NT) k_up3(const ConvArgs a, const StripPlan p) {
    constexpr int KK = K * K;
    extern __shared__ __align__(16) float s_w[];       // [ci][tap][COT]
    const int tid = threadIdx.x;
    const int co0 = blockIdx.y * COT;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    const int Hin = iv.H, Win = iv.W, Hout = a.out.H;
    const int QH = p.RP;
    for (int i = tid; i < a.Cin * KK * COT; i += CAE_NT) {
        int j = i % COT, t = (i / COT) % KK, ci = i / (COT * KK);
        int co = co0 + j;
        s_w[i] = co < a.Cout ? __ldg(a.w + ((size_t)ci * a.Cout + co) * KK + t) : 0.f;
    }
    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;
    __syncthreads();

    for (int u0 = blockIdx.x * CAE_NT; u0 < p.units; u0 += gridDim.x * CAE_NT) {
        const int u = u0 + tid;
        if (u < p.units) {
            const int R = u / p.NS, st = u - R * p.NS;
            const int n = R / QH, qy = R - n * QH;
            const int qx0 = 4 * st;
            float acc[COT][2][8];
#pragma unroll
            for (int j = 0; j < COT; ++j)
#pragma unroll
                for (int py = 0; py < 2; ++py)
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[j][py][e] = 0.f;

            const long long nbase = in_base + (long long)n * iv.sN;
            const bool vec_ok = qx0 < Win;            // the aligned float4 starts inside the row (rows are padded to 4)
            for (int ci = 0; ci < a.Cin; ++ci) {
                const ChanCoef kc = load_coef(a.in, ci);
                if (ci + 1 < a.Cin && vec_ok) {
#pragma unroll
                    for (int jy = 0; jy < 2; ++jy) {
                        const int iy = qy - jy;
                        if (iy >= 0 && iy < Hin) pf_src(a.in, nbase + (long long)(ci + 1) * iv.sC + (long long)iy * iv.ld + qx0);
                    }
                }
                float v[2][5];                         // [jy][x - (qx0-1)]
#pragma unroll
                for (int jy = 0; jy < 2; ++jy) {
                    const int iy = qy - jy;
                    const bool rok = iy >= 0 && iy < Hin;
                    const long long ro = nbase + (long long)ci * iv.sC + (long long)iy * iv.ld + qx0;
                    float t4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (rok && vec_ok) {
                        load4(a.in, ro, kc, t4);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (qx0 + i >= Win) t4[i] = 0.f;
                    }
                    v[jy][0] = (rok && qx0 > 0 && qx0 - 1 < Win) ? src_value(a.in, ro - 1, kc) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[jy][1 + i] = t4[i];
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
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const int oy = 2 * qy + py;
                if (oy < Hout) {
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < a.Cout)
                            epi_strip<8>(a.epi, a.out, ech[j], n, co0 + j, oy, 2 * qx0, acc[j][py], tgt_base, a.inv_count,
                                         s1[j], s2[j]);
                }
            }
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}

// =======================================================================================
// DOWN v3: strided conv, stride 2, pad 0.  Thread = 4 output pixels x COT channels.
// =======================================================================================
template <int K, int COT>
__global__ void __launch_bounds__(CAE_NT) k_down3(const ConvArgs a, const StripPlan p) {
    constexpr int KK = K * K;
    constexpr int NVV = 8 + K - 2;                    // input columns needed: 2*ox0 .. 2*ox0 + 6 + K - 1
    extern __shared__ __align__(16) float s_w[];      // [ci][tap][COT]
    const int tid = threadIdx.x;
    const int co0 = blockIdx.y * COT;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const int Hin = iv.H, Win = iv.W;
    const int OH = p.RP;
    for (int i = tid; i < a.Cin * KK * COT; i += CAE_NT) {
        int j = i % COT, t = (i / COT) % KK, ci = i / (COT * KK);
        int co = co0 + j;
        s_w[i] = co < a.Cout ? __ldg(a.w + ((size_t)co * a.Cin + ci) * KK + t) : 0.f;
    }
    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;
    __syncthreads();

    for (int u0 = blockIdx.x * CAE_NT; u0 < p.units; u0 += gridDim.x * CAE_NT) {
        const int u = u0 + tid;
        if (u < p.units) {
            const int R = u / p.NS, st = u - R * p.NS;
            const int n = R / OH, oy = R - n * OH;
            const int ox0 = 4 * st, xb = 8 * st;
            float acc[COT][4];
#pragma unroll
            for (int j = 0; j < COT; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
            const long long nbase = in_base + (long long)n * iv.sN + xb;
            for (int ci = 0; ci < a.Cin; ++ci) {
                const ChanCoef kc = load_coef(a.in, ci);
                const float* wp = s_w + ci * KK * COT;
                if (ci + 1 < a.Cin && xb < Win) {
#pragma unroll
                    for (int ky = 0; ky < K; ++ky)
                        if (2 * oy + ky < Hin) pf_src(a.in, nbase + (long long)(ci + 1) * iv.sC + (long long)(2 * oy + ky) * iv.ld);
                }
#pragma unroll
                for (int ky = 0; ky < K; ++ky) {
                    const int r = 2 * oy + ky;
                    const bool rok = r < Hin;
                    const long long ro = nbase + (long long)ci * iv.sC + (long long)r * iv.ld;
                    float v[NVV + 3];
#pragma unroll
                    for (int i = 0; i < NVV + 3; ++i) v[i] = 0.f;
                    if (rok) {
                        float t4[4];
                        if (xb < Win) {
                            load4(a.in, ro, kc, t4);
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = (xb + i < Win) ? t4[i] : 0.f;
                        }
                        if (xb + 4 < Win) {
                            load4(a.in, ro + 4, kc, t4);
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[4 + i] = (xb + 4 + i < Win) ? t4[i] : 0.f;
                        }
#pragma unroll
                        for (int i = 8; i < NVV; ++i) v[i] = (xb + i < Win) ? src_value(a.in, ro + i, kc) : 0.f;
                    }
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        float wv[COT];
#pragma unroll
                        for (int j = 0; j < COT; ++j) wv[j] = wp[(ky * K + kx) * COT + j];
#pragma unroll
                        for (int cx = 0; cx < 4; ++cx)
#pragma unroll
                            for (int j = 0; j < COT; ++j) acc[j][cx] = fmaf(v[2 * cx + kx], wv[j], acc[j][cx]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < a.Cout)
                    epi_strip<4>(a.epi, a.out, ech[j], n, co0 + j, oy, ox0, acc[j], 0ll, a.inv_count, s1[j], s2[j]);
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}

// =======================================================================================
// WGRAD v3 (wide thin layers, stride 2, pad 0):
//   G[cs][cb][ky][kx] = sum_{n,i,j} S(n,cs,i,j) * B(n,cb,2i+ky,2j+kx)
// Thread = strip of 4 positions of S (float4) and the matching 2x-strided window of B (two float4 + tail);
// a CST x CBT x K*K register tile of partial sums lives in the thread for the whole (persistent) kernel; then
// warp shuffle -> shared memory -> one partial row per CTA -> fixed-order sum by the last CTA.
// blockIdx.y enumerates the (cs tile, cb tile) pairs.
// =======================================================================================
template <int K, int CST, int CBT>
__global__ void __launch_bounds__(CAE_NT) k_wgrad3(const WgradArgs a, const StripPlan p, int tiles_b) {
    constexpr int KK = K * K;
    constexpr int NVV = 8 + K - 2;
    constexpr int NACC = CST * CBT * KK;
    __shared__ float red[CAE_NWARP][NACC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cs0 = (blockIdx.y / tiles_b) * CST, cb0 = (blockIdx.y % tiles_b) * CBT;
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    const int Hs = p.RP, Ws = sv.W, Hb = bv.H, Wb = bv.W;

    ChanCoef ks[CST], kb[CBT];
#pragma unroll
    for (int x = 0; x < CST; ++x) ks[x] = load_coef(a.sm, min(cs0 + x, a.Cs - 1));
#pragma unroll
    for (int y = 0; y < CBT; ++y) kb[y] = load_coef(a.bg, min(cb0 + y, a.Cb - 1));

    float acc[CST][CBT][KK];
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) acc[x][y][t] = 0.f;

    for (int u0 = blockIdx.x * CAE_NT; u0 < p.units; u0 += gridDim.x * CAE_NT) {
        const int u = u0 + tid;
        if (u < p.units) {
            const int R = u / p.NS, st = u - R * p.NS;
            const int n = R / Hs, i = R - n * Hs;
            const int j0 = 4 * st, xb = 8 * st;
            float sval[CST][4];
#pragma unroll
            for (int x = 0; x < CST; ++x) {
                float t4[4] = {0.f, 0.f, 0.f, 0.f};
                if (cs0 + x < a.Cs) {
                    load4(a.sm, sbase + (long long)n * sv.sN + (long long)(cs0 + x) * sv.sC + (long long)i * sv.ld + j0, ks[x], t4);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (j0 + e >= Ws) t4[e] = 0.f;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) sval[x][e] = t4[e];
            }
#pragma unroll
            for (int y = 0; y < CBT; ++y) {
                if (cb0 + y < a.Cb) {
                    const long long cbase = bbase + (long long)n * bv.sN + (long long)(cb0 + y) * bv.sC + xb;
#pragma unroll
                    for (int ky = 0; ky < K; ++ky) {
                        const int r = 2 * i + ky;
                        float v[NVV + 3];
#pragma unroll
                        for (int q = 0; q < NVV + 3; ++q) v[q] = 0.f;
                        if (r < Hb) {
                            const long long ro = cbase + (long long)r * bv.ld;
                            float t4[4];
                            if (xb < Wb) {
                                load4(a.bg, ro, kb[y], t4);
#pragma unroll
                                for (int q = 0; q < 4; ++q) v[q] = (xb + q < Wb) ? t4[q] : 0.f;
                            }
                            if (xb + 4 < Wb) {
                                load4(a.bg, ro + 4, kb[y], t4);
#pragma unroll
                                for (int q = 0; q < 4; ++q) v[4 + q] = (xb + 4 + q < Wb) ? t4[q] : 0.f;
                            }
#pragma unroll
                            for (int q = 8; q < NVV; ++q) v[q] = (xb + q < Wb) ? src_value(a.bg, ro + q, kb[y]) : 0.f;
                        }
#pragma unroll
                        for (int kx = 0; kx < K; ++kx)
#pragma unroll
                            for (int e = 0; e < 4; ++e)
#pragma unroll
                                for (int x = 0; x < CST; ++x)
                                    acc[x][y][ky * K + kx] = fmaf(sval[x][e], v[2 * e + kx], acc[x][y][ky * K + kx]);
                    }
                }
            }
        }
    }
    // lanes -> warp -> CTA
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                float s = warp_sum(acc[x][y][t]);
                if (lane == 0) red[warp][(x * CBT + y) * KK + t] = s;
            }
    __syncthreads();
    const int nelem = a.Cs * a.Cb * KK;
    for (int i = tid; i < NACC; i += CAE_NT) {
        const int x = i / (CBT * KK), y = (i / KK) % CBT, t = i % KK;
        if (cs0 + x < a.Cs && cb0 + y < a.Cb) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < CAE_NWARP; ++w) s += red[w][i];
            a.partials[(size_t)blockIdx.x * nelem + ((size_t)(cs0 + x) * a.Cb + (cb0 + y)) * KK + t] = s;
        }
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
