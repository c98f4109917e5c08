// v2 conv kernels for stride 2 with 3/4-wide kernels: shared-memory staged input tiles (on-load transform
// applied once per element), register tiles of CX consecutive positions x COT channels, 32-bit indexing,
// rows of all samples flattened into one axis so that layers with tiny spatial extent still fill CTAs.
#pragma once
#include "conv_family.cuh"

// COT consecutive floats from shared memory (16-byte aligned when COT % 4 == 0)
template <int COT>
__device__ __forceinline__ void load_wvec(float (&wv)[COT], const float* p) {
    if constexpr (COT % 4 == 0) {
#pragma unroll
        for (int q = 0; q < COT / 4; ++q) {
            float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
        }
    } else if constexpr (COT == 2) {
        float2 t = *reinterpret_cast<const float2*>(p);
        wv[0] = t.x; wv[1] = t.y;
    } else {
#pragma unroll
        for (int j = 0; j < COT; ++j) wv[j] = p[j];
    }
}

struct TilePlan {
    int TXT;        // threads along x per tile row (power of two, <= 32); TYT = CAE_NT / TXT
    int txt_shift;  // log2(TXT)
    int RP;         // rows per sample in the flattened row axis (cells rows / padded output rows)
    int total_rows; // N * RP
    int ncol_tiles, nrow_tiles;
    int SROWS;      // staged rows per channel
    int SCP;        // staged row pitch (floats, multiple of 4)
    int ci_chunk;   // input channels staged per pass
    int TZ;         // thread groups along the output-channel axis inside a CTA (each owns COT channels)
    int tyt_shift;  // log2(TYT): tid = (((tk * TZ + tz) * TYT) + ty) * TXT + tx
    int TK;         // split of the input-channel reduction over thread groups (partial sums added through smem)
    int tz_shift;   // log2(TZ)
};

// Add the TK partial accumulators of every output through shared memory (fixed order); the result is left in
// the threads with tk == 0.  scratch: >= blockDim.x * NACC floats.
template <int NACC>
__device__ __forceinline__ void reduce_over_tk(float* acc, float* scratch, int tk, int TK) {
    const int slice = blockDim.x / TK;
    __syncthreads();
    if (tk > 0) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) scratch[(size_t)i * blockDim.x + threadIdx.x] = acc[i];
    }
    __syncthreads();
    if (tk == 0) {
        for (int k = 1; k < TK; ++k) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += scratch[(size_t)i * blockDim.x + threadIdx.x + k * slice];
        }
    }
}

// ---------------------------------------------------------------------------------------
// epilogue for a run of NE consecutive output pixels in one row (32-bit offsets)
// ---------------------------------------------------------------------------------------
template <int NE>
__device__ __forceinline__ void epi_row(const CaeEpilogue& e, const CaeView& out, const EpiCh& ch, int n, int co, int oy,
                                        int ox0, const float (&acc_in)[NE], long long tgt_base, float inv_count, float& s1,
                                        float& s2) {
    const int Wout = out.W;
    float* orow = out.p + ((long long)n * out.sN + (long long)co * out.sC + (long long)oy * out.ld);
    float accl[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) accl[i] = acc_in[i];
    if (e.addend.t0.p) {                  // second gradient of the same geometry (skip-connection fan-in)
        const CaeView& av = e.addend.t0;
        const long long aoff = (long long)n * av.sN + (long long)co * av.sC + (long long)oy * av.ld;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            int ox = ox0 + i;
            if (ox >= 0 && ox < Wout) accl[i] += src_value(e.addend, aoff + ox, ch.ak);
        }
    }
    const float (&acc)[NE] = accl;
    switch (e.mode) {
        case CAE_EPI_MASK: {
            const CaeView& a = e.act;
            const float* arow = a.p + ((long long)n * a.sN + (long long)co * a.sC + (long long)oy * a.ld);
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int ox = ox0 + i;
                if (ox >= 0 && ox < Wout) orow[ox] = __ldg(arow + ox) > 0.f ? acc[i] : 0.f;
            }
        } break;
        case CAE_EPI_PLAIN:
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int ox = ox0 + i;
                if (ox >= 0 && ox < Wout) orow[ox] = acc[i] + ch.bias;
            }
            break;
        case CAE_EPI_STATS:
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int ox = ox0 + i;
                if (ox >= 0 && ox < Wout) {
                    float v = acc[i] + ch.bias;
                    orow[ox] = v;
                    s1 += v;
                    s2 = fmaf(v, v, s2);
                }
            }
            break;
        case CAE_EPI_MASKSTATS: {
            const CaeView& a = e.act;
            const float* arow = a.p + ((long long)n * a.sN + (long long)co * a.sC + (long long)oy * a.ld);
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int ox = ox0 + i;
                if (ox >= 0 && ox < Wout) {
                    float yp = __ldg(arow + ox);
                    float z = fmaf(yp, ch.scale, ch.shift);
                    float dz = z > 0.f ? acc[i] : 0.f;
                    orow[ox] = dz;
                    s1 += dz;
                    s2 = fmaf(dz, (yp - ch.mean) * ch.invstd, s2);
                }
            }
        } break;
        case CAE_EPI_SIGMOID:
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int ox = ox0 + i;
                if (ox >= 0 && ox < Wout) orow[ox] = 1.f / (1.f + expf(-(acc[i] + ch.bias)));
            }
            break;
        case CAE_EPI_SIGMOID_MSE: {
            const CaeView& t = e.target.t0;
            const long long toff = tgt_base + (long long)n * t.sN + (long long)co * t.sC + (long long)oy * t.ld;
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int ox = ox0 + i;
                if (ox >= 0 && ox < Wout) {
                    float yh = 1.f / (1.f + expf(-(acc[i] + ch.bias)));
                    float d = yh - src_value(e.target, toff + ox, ch.tk);
                    s2 = fmaf(d, d, s2);
                    float dz = 2.f * d * inv_count * yh * (1.f - yh);
                    s1 += dz;
                    if (e.write_mode == 0) orow[ox] = dz;
                    else if (e.write_mode == 1) orow[ox] = yh;
                }
            }
        } break;
    }
}

// CTA-level reduction when the CTA's threads are split into TZ channel groups (group tz = tid / gthreads owns
// channels cbase + tz*COT ..): every thread parks its sums in shared memory (`scratch`, >= blockDim * 2*COT
// floats, the tile buffers are free by now), one thread per (channel, statistic) adds the group's entries in
// thread order (deterministic) into this CTA's partial row; then the usual last-CTA finalize.
template <int COT>
__device__ __forceinline__ void epi_reduce_tail_tz(const CaeEpilogue& e, const CaeView& out, float* scratch, int cbase,
                                                   int WC, int gthreads, const float (&s1)[COT], const float (&s2)[COT]) {
    const int C = out.C;
    __syncthreads();
    if (gthreads >= 32) {
        // every warp lies inside one channel group: shuffle-reduce, then add the group's warps in order
        double* dscr = reinterpret_cast<double*>(scratch);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < COT; ++j) {
            double a = warp_sum_d((double)s1[j]), b = warp_sum_d((double)s2[j]);
            if (lane == 0) {
                dscr[warp * (2 * COT) + 2 * j] = a;
                dscr[warp * (2 * COT) + 2 * j + 1] = b;
            }
        }
        __syncthreads();
        const int wpg = gthreads >> 5;
        for (int i = threadIdx.x; i < WC * 2; i += blockDim.x) {
            const int ch = i >> 1, g = ch / COT, j = ch - g * COT, st = i & 1;
            const int co = cbase + ch;
            if (co < C) {
                double s = 0.0;
                for (int w = 0; w < wpg; ++w) s += dscr[(g * wpg + w) * (2 * COT) + 2 * j + st];
                e.partials[((size_t)blockIdx.x * C + co) * 2 + st] = s;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < COT; ++j) {
            scratch[threadIdx.x * (2 * COT) + 2 * j] = s1[j];
            scratch[threadIdx.x * (2 * COT) + 2 * j + 1] = s2[j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < WC * 2; i += blockDim.x) {
            const int ch = i >> 1, g = ch / COT, j = ch - g * COT, st = i & 1;
            const int co = cbase + ch;
            if (co < C) {
                double s = 0.0;
                const float* src = scratch + (size_t)g * gthreads * (2 * COT) + 2 * j + st;
                for (int t = 0; t < gthreads; ++t) s += (double)src[t * (2 * COT)];
                e.partials[((size_t)blockIdx.x * C + co) * 2 + st] = s;
            }
        }
    }
    if (cae_last_block(e.ticket)) {
        const double count = (double)out.N * out.H * out.W;
        if (e.mode == CAE_EPI_STATS) finalize_bn_forward(e.bn, e.partials, gridDim.x, count);
        else if (e.mode == CAE_EPI_MASKSTATS) finalize_bn_backward(e.bn, e.partials, gridDim.x, count);
        else finalize_mse(e, e.partials, gridDim.x, C, count * C);
    }
}

// =======================================================================================
// UP v2 (transposed conv, gather form).  Thread = CX consecutive cells x COT channels.
// Flattened row axis: R = n * QH + qy, QH cell rows per sample (QH - Hin >= JY-1 zero rows
// separate the samples, so a tap that falls off the top of a sample reads a staged zero row).
// =======================================================================================
template <int KH, int KW, int CX, int COT>
__global__ void __launch_bounds__(CAE_NT) k_up2(const ConvArgs a, const TilePlan p) {
    constexpr int JY = (KH + 1) / 2, JX = (KW + 1) / 2, KK = KH * KW, NV = CX + JX - 1;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;
    float* s_w = smem + p.ci_chunk * p.SROWS * p.SCP;
    long long* s_row = reinterpret_cast<long long*>(s_w + ((p.ci_chunk * KK * COT * p.TZ + 3) & ~3));   // [SROWS] row offsets
    const int tid = threadIdx.x;
    const int TYT = 1 << p.tyt_shift;
    const int tx = tid & (p.TXT - 1), ty = (tid >> p.txt_shift) & (TYT - 1);
    const int tz = (tid >> (p.txt_shift + p.tyt_shift)) & (p.TZ - 1), tk = tid >> (p.txt_shift + p.tyt_shift + p.tz_shift);
    const int WC = COT * p.TZ;                      // channels per CTA
    const int co0 = blockIdx.y * WC + tz * COT;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    const int Hin = iv.H, Win = iv.W, Hout = a.out.H;
    const int QH = p.RP;
    const int SCOLS = p.TXT * CX + JX - 1;

    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;

    const int ntiles = p.nrow_tiles * p.ncol_tiles;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int rt = tile / p.ncol_tiles, ct = tile - rt * p.ncol_tiles;
        const int R0 = rt * TYT, qx0 = ct * p.TXT * CX;
        const int R = R0 + ty;
        const bool rowvalid = R < p.total_rows;
        const int n = rowvalid ? R / QH : 0;
        const int qy = R - n * QH;

        float acc[COT][2][2 * CX];
#pragma unroll
        for (int j = 0; j < COT; ++j)
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int e = 0; e < 2 * CX; ++e) acc[j][py][e] = 0.f;

        // offsets of the staged rows (channel 0), -1 for rows that read as zero
        __syncthreads();
        for (int u = tid; u < p.SROWS; u += blockDim.x) {
            const int Rr = R0 - (JY - 1) + u;
            long long off = -1;
            if (Rr >= 0 && Rr < p.total_rows) {
                const int nn = Rr / QH, q = Rr - nn * QH;
                if (q < Hin) off = in_base + (long long)nn * iv.sN + (long long)q * iv.ld;
            }
            s_row[u] = off;
        }
        for (int c0 = 0; c0 < a.Cin; c0 += p.ci_chunk) {
            const int cn = min(p.ci_chunk, a.Cin - c0);
            __syncthreads();
            // weights of this chunk: [cl][tap][WC]
            for (int i = tid; i < cn * KK * WC; i += blockDim.x) {
                int j = i % WC, t = (i / WC) % KK, cl = i / (WC * KK);
                int co = blockIdx.y * WC + j;
                s_w[i] = co < a.Cout ? __ldg(a.w + ((size_t)(c0 + cl) * a.Cout + co) * KK + t) : 0.f;
            }
            // inputs: element-parallel so narrow rows still keep every lane busy
            {
                const int per_ch = p.SROWS * SCOLS, total = cn * per_ch;
#pragma unroll 4
                for (int idx = tid; idx < total; idx += blockDim.x) {
                    const int cl = idx / per_ch, rem = idx - cl * per_ch;
                    const int u = rem / SCOLS, v = rem - u * SCOLS;
                    const int ix = qx0 - (JX - 1) + v;
                    const long long off = s_row[u];
                    float val = 0.f;
                    if (off >= 0 && ix >= 0 && ix < Win) {
                        const ChanCoef kc = load_coef(a.in, c0 + cl);
                        val = src_value(a.in, off + (long long)(c0 + cl) * iv.sC + ix, kc);
                    }
                    s_in[(cl * p.SROWS + u) * p.SCP + v] = val;
                }
            }
            __syncthreads();
#pragma unroll 2
            for (int cl = tk; cl < cn; cl += p.TK) {
                float v[JY][NV];
#pragma unroll
                for (int jy = 0; jy < JY; ++jy) {
                    const float* sp = s_in + (cl * p.SROWS + ty + (JY - 1) - jy) * p.SCP + tx * CX;
                    if constexpr (CX == 4) {
                        float4 q4 = *reinterpret_cast<const float4*>(sp);
                        v[jy][0] = q4.x; v[jy][1] = q4.y; v[jy][2] = q4.z; v[jy][3] = q4.w;
#pragma unroll
                        for (int i = 4; i < NV; ++i) v[jy][i] = sp[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < NV; ++i) v[jy][i] = sp[i];
                    }
                }
                const float* wp = s_w + cl * KK * WC + tz * COT;
#pragma unroll
                for (int py = 0; py < 2; ++py)
#pragma unroll
                    for (int jy = 0; jy < JY; ++jy) {
                        const int ky = py + 2 * jy;
                        if (ky < KH) {
#pragma unroll
                            for (int px = 0; px < 2; ++px)
#pragma unroll
                                for (int jx = 0; jx < JX; ++jx) {
                                    const int kx = px + 2 * jx;
                                    if (kx < KW) {
                                        float wv[COT];
                                        load_wvec<COT>(wv, wp + (ky * KW + kx) * WC);
#pragma unroll
                                        for (int cx = 0; cx < CX; ++cx) {
                                            const float xv = v[jy][cx + (JX - 1) - jx];
#pragma unroll
                                            for (int j = 0; j < COT; ++j)
                                                acc[j][py][2 * cx + px] = fmaf(xv, wv[j], acc[j][py][2 * cx + px]);
                                        }
                                    }
                                }
                        }
                    }
            }
        }
        if (p.TK > 1) reduce_over_tk<COT * 4 * CX>(&acc[0][0][0], smem, tk, p.TK);
        // epilogue
        if (rowvalid && tk == 0) {
            const int ox0 = 2 * (qx0 + tx * CX) - a.p;
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const int oy = 2 * qy + py - a.p;
                if (oy >= 0 && oy < Hout) {
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < a.Cout)
                            epi_row<2 * CX>(a.epi, a.out, ech[j], n, co0 + j, oy, ox0, acc[j][py], tgt_base, a.inv_count,
                                            s1[j], s2[j]);
                }
            }
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail_tz<COT>(a.epi, a.out, smem, blockIdx.y * WC, WC, TYT * p.TXT, s1, s2);
}

// =======================================================================================
// DOWN v2 (strided conv).  Thread = CX consecutive output pixels x COT channels.
// Flattened output rows: R = n * OHp + oy with OHp = Hout + 1; staged input rows are the flattened
// input rows 2*R0 .. 2*(R0+TYT-1)+KH-1 (input row of sample n: r - pad, r = F - n*2*OHp).
// Columns are staged de-interleaved (even / odd taps) so that stride-2 reads are conflict free:
//   A0[c] = in[2c - pad], A1[c] = in[2c + 1 - pad];  tap kx of output ox reads A(kx&1)[ox + (kx>>1)].
// =======================================================================================
template <int KH, int KW, int CX, int COT>
__global__ void __launch_bounds__(CAE_NT) k_down2(const ConvArgs a, const TilePlan p) {
    constexpr int KK = KH * KW;
    constexpr int NV0 = CX + (KW - 1) / 2;   // even taps: kx = 0, 2
    constexpr int NV1 = CX + (KW - 2) / 2;   // odd taps:  kx = 1, (3)
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                                           // [cl][SROWS][2][SCP]
    float* s_w = smem + p.ci_chunk * p.SROWS * 2 * p.SCP;         // [cl][tap][WC]
    long long* s_row = reinterpret_cast<long long*>(s_w + ((p.ci_chunk * KK * COT * p.TZ + 3) & ~3));
    const int tid = threadIdx.x;
    const int TYT = 1 << p.tyt_shift;
    const int tx = tid & (p.TXT - 1), ty = (tid >> p.txt_shift) & (TYT - 1);
    const int tz = (tid >> (p.txt_shift + p.tyt_shift)) & (p.TZ - 1), tk = tid >> (p.txt_shift + p.tyt_shift + p.tz_shift);
    const int WC = COT * p.TZ;
    const int co0 = blockIdx.y * WC + tz * COT;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const int Hin = iv.H, Win = iv.W, Hout = a.out.H;
    const int OHp = p.RP;
    const int NC = p.TXT * CX + 2;   // staged columns per parity

    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;

    const int ntiles = p.nrow_tiles * p.ncol_tiles;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int rt = tile / p.ncol_tiles, ct = tile - rt * p.ncol_tiles;
        const int R0 = rt * TYT, ox0t = ct * p.TXT * CX;
        const int R = R0 + ty;
        const bool rowvalid = R < p.total_rows;
        const int n = rowvalid ? R / OHp : 0;
        const int oy = R - n * OHp;

        float acc[COT][CX];
#pragma unroll
        for (int j = 0; j < COT; ++j)
#pragma unroll
            for (int e = 0; e < CX; ++e) acc[j][e] = 0.f;

        __syncthreads();
        for (int w = tid; w < p.SROWS; w += blockDim.x) {
            const int F = 2 * R0 + w;                       // flattened input row
            const int nn = F / (2 * OHp);
            const int r = F - nn * 2 * OHp - a.p;           // input row inside the sample
            s_row[w] = (nn < a.out.N && r >= 0 && r < Hin) ? in_base + (long long)nn * iv.sN + (long long)r * iv.ld : -1;
        }
        for (int c0 = 0; c0 < a.Cin; c0 += p.ci_chunk) {
            const int cn = min(p.ci_chunk, a.Cin - c0);
            __syncthreads();
            for (int i = tid; i < cn * KK * WC; i += blockDim.x) {
                int j = i % WC, t = (i / WC) % KK, cl = i / (WC * KK);
                int co = blockIdx.y * WC + j;
                s_w[i] = co < a.Cout ? __ldg(a.w + ((size_t)co * a.Cin + (c0 + cl)) * KK + t) : 0.f;
            }
            {
                const int rowlen = 2 * NC, per_ch = p.SROWS * rowlen, total = cn * per_ch;
                const int xb = 2 * ox0t - a.p;                  // input column of A0[0]
#pragma unroll 4
                for (int idx = tid; idx < total; idx += blockDim.x) {
                    const int cl = idx / per_ch, rem = idx - cl * per_ch;
                    const int w = rem / rowlen, v = rem - w * rowlen;
                    const int ix = xb + v;
                    const long long off = s_row[w];
                    float val = 0.f;
                    if (off >= 0 && ix >= 0 && ix < Win) {
                        const ChanCoef kc = load_coef(a.in, c0 + cl);
                        val = src_value(a.in, off + (long long)(c0 + cl) * iv.sC + ix, kc);
                    }
                    s_in[((cl * p.SROWS + w) * 2 + (v & 1)) * p.SCP + (v >> 1)] = val;
                }
            }
            __syncthreads();
#pragma unroll 2
            for (int cl = tk; cl < cn; cl += p.TK) {
                const float* wp = s_w + cl * KK * WC + tz * COT;
#pragma unroll
                for (int ky = 0; ky < KH; ++ky) {
                    const float* sp = s_in + (cl * p.SROWS + 2 * ty + ky) * 2 * p.SCP + tx * CX;
                    float e0[NV0], e1[NV1];
                    if constexpr (CX == 4) {
                        float4 q0 = *reinterpret_cast<const float4*>(sp);
                        float4 q1 = *reinterpret_cast<const float4*>(sp + p.SCP);
                        e0[0] = q0.x; e0[1] = q0.y; e0[2] = q0.z; e0[3] = q0.w;
                        e1[0] = q1.x; e1[1] = q1.y; e1[2] = q1.z; e1[3] = q1.w;
#pragma unroll
                        for (int i = 4; i < NV0; ++i) e0[i] = sp[i];
#pragma unroll
                        for (int i = 4; i < NV1; ++i) e1[i] = sp[p.SCP + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < NV0; ++i) e0[i] = sp[i];
#pragma unroll
                        for (int i = 0; i < NV1; ++i) e1[i] = sp[p.SCP + i];
                    }
#pragma unroll
                    for (int kx = 0; kx < KW; ++kx) {
                        float wv[COT];
                        load_wvec<COT>(wv, wp + (ky * KW + kx) * WC);
#pragma unroll
                        for (int cx = 0; cx < CX; ++cx) {
                            const float xv = (kx & 1) ? e1[cx + (kx >> 1)] : e0[cx + (kx >> 1)];
#pragma unroll
                            for (int j = 0; j < COT; ++j) acc[j][cx] = fmaf(xv, wv[j], acc[j][cx]);
                        }
                    }
                }
            }
        }
        if (p.TK > 1) reduce_over_tk<COT * CX>(&acc[0][0], smem, tk, p.TK);
        if (rowvalid && oy < Hout && tk == 0) {
            const int ox0 = ox0t + tx * CX;
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < a.Cout)
                    epi_row<CX>(a.epi, a.out, ech[j], n, co0 + j, oy, ox0, acc[j], 0ll, a.inv_count, s1[j], s2[j]);
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail_tz<COT>(a.epi, a.out, smem, blockIdx.y * WC, WC, TYT * p.TXT, s1, s2);
}

// =======================================================================================
// WGRAD v2a - "position parallel" (few channels, large spatial extent).
//   G[cs][cb][ky][kx] = sum_{n,i,j} S(n,cs,i,j) * B(n,cb,2i+ky-p,2j+kx-p)
// A CTA walks tiles of TR flattened rows (R = n*(Hs+1) + i); both operand tiles are staged in shared
// memory (B de-interleaved like k_down2).  Warp w owns the (cs,cb) register tile g = w % GP and the tile
// rows r = w / GP, w / GP + 8 / GP, ...; lanes own CX consecutive positions of the row.  After the last
// tile: warp-shuffle reduction, cross-warp sum in shared memory, one partial row per CTA; the last CTA
// of the grid sums the rows in index order.
// =======================================================================================
struct Wg2Plan {
    int RP, total_rows;   // flattened rows: Hs + 1 per sample
    int TR;               // tile rows
    int nrow_tiles, ncol_tiles;
    int SCPs, SCPb;       // row pitches of the S tile and (per parity) of the B tile
    int tiles_b;          // ceil(Cb / CBT)
    int G, GP;            // register-tile groups, padded to a power of two (<= 8)
};

template <int KH, int KW, int CX, int CST, int CBT>
__global__ void __launch_bounds__(CAE_NT) k_wgrad2a(const WgradArgs a, const Wg2Plan p) {
    constexpr int KK = KH * KW;
    constexpr int NV0 = CX + (KW - 1) / 2, NV1 = CX + (KW - 2) / 2;
    constexpr int NACC = CST * CBT * KK;
    extern __shared__ __align__(16) float smem[];
    const int BROWS = 2 * p.TR + KH - 2;
    float* s_s = smem;                                   // [Cs][TR][SCPs]
    float* s_b = smem + a.Cs * p.TR * p.SCPs;            // [Cb][BROWS][2][SCPb]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    const int Hs = sv.H, Ws = sv.W, Hb = bv.H, Wb = bv.W, N = sv.N;
    const int g = warp % p.GP, wrow = warp / p.GP, wstep = CAE_NWARP / p.GP;
    const bool gactive = g < p.G;
    const int cs0 = (g / p.tiles_b) * CST, cb0 = (g % p.tiles_b) * CBT;
    const int TXC = 32 * CX;
    const int NCb = TXC + 2;

    float acc[CST][CBT][KK];
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) acc[x][y][t] = 0.f;

    const int ntiles = p.nrow_tiles * p.ncol_tiles;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int rt = tile / p.ncol_tiles, ct = tile - rt * p.ncol_tiles;
        const int R0 = rt * p.TR, j0 = ct * TXC;
        __syncthreads();
        // stage S rows
        for (int task = warp; task < a.Cs * p.TR; task += CAE_NWARP) {
            const int c = task / p.TR, r = task - c * p.TR;
            const int R = R0 + r;
            int n = 0, i = 0;
            bool ok = R < p.total_rows;
            if (ok) {
                n = R / p.RP;
                i = R - n * p.RP;
                ok = i < Hs;
            }
            const ChanCoef kc = load_coef(a.sm, c);
            const long long rb = sbase + (long long)n * sv.sN + (long long)c * sv.sC + (long long)i * sv.ld;
            float* dst = s_s + (c * p.TR + r) * p.SCPs;
            for (int v = lane; v < TXC; v += 32) {
                const int j = j0 + v;
                dst[v] = (ok && j < Ws) ? src_value(a.sm, rb + j, kc) : 0.f;
            }
        }
        // stage B rows (de-interleaved)
        for (int task = warp; task < a.Cb * BROWS; task += CAE_NWARP) {
            const int c = task / BROWS, w = task - c * BROWS;
            const int F = 2 * R0 + w;
            const int n = F / (2 * p.RP);
            const int r = F - n * 2 * p.RP - a.p;
            const bool ok = n < N && r >= 0 && r < Hb;
            const ChanCoef kc = load_coef(a.bg, c);
            const long long rb = bbase + (long long)n * bv.sN + (long long)c * bv.sC + (long long)r * bv.ld;
            float* dst = s_b + (c * BROWS + w) * 2 * p.SCPb;
            const int xb = 2 * j0 - a.p;
            for (int v = lane; v < 2 * NCb; v += 32) {
                const int ix = xb + v;
                float val = (ok && ix >= 0 && ix < Wb) ? src_value(a.bg, rb + ix, kc) : 0.f;
                dst[(v & 1) * p.SCPb + (v >> 1)] = val;
            }
        }
        __syncthreads();
        if (gactive) {
            for (int r = wrow; r < p.TR; r += wstep) {
                float sval[CST][CX];
#pragma unroll
                for (int x = 0; x < CST; ++x) {
                    const int cs = min(cs0 + x, a.Cs - 1);
                    const float* sp = s_s + (cs * p.TR + r) * p.SCPs + lane * CX;
                    const bool live = cs0 + x < a.Cs;
#pragma unroll
                    for (int e = 0; e < CX; ++e) sval[x][e] = live ? sp[e] : 0.f;
                }
#pragma unroll
                for (int y = 0; y < CBT; ++y) {
                    const int cb = min(cb0 + y, a.Cb - 1);
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky) {
                        const float* bp = s_b + (cb * BROWS + 2 * r + ky) * 2 * p.SCPb + lane * CX;
                        float e0[NV0], e1[NV1];
#pragma unroll
                        for (int i = 0; i < NV0; ++i) e0[i] = bp[i];
#pragma unroll
                        for (int i = 0; i < NV1; ++i) e1[i] = bp[p.SCPb + i];
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx)
#pragma unroll
                            for (int e = 0; e < CX; ++e) {
                                const float b = (kx & 1) ? e1[e + (kx >> 1)] : e0[e + (kx >> 1)];
#pragma unroll
                                for (int x = 0; x < CST; ++x)
                                    acc[x][y][ky * KW + kx] = fmaf(sval[x][e], b, acc[x][y][ky * KW + kx]);
                            }
                    }
                }
            }
        }
    }
    // ---- reduce: lanes -> warp, warps of the same group -> CTA partial row
    __syncthreads();
    float* red = smem;   // [warp][NACC] (tiles are no longer needed)
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                float v = warp_sum(acc[x][y][t]);
                if (lane == 0) red[warp * NACC + (x * CBT + y) * KK + t] = v;
            }
    __syncthreads();
    const int nelem = a.Cs * a.Cb * KK;
    for (int i = tid; i < p.G * NACC; i += CAE_NT) {
        const int gg = i / NACC, rem = i - gg * NACC;
        const int x = rem / (CBT * KK), y = (rem / KK) % CBT, t = rem % KK;
        const int cs = (gg / p.tiles_b) * CST + x, cb = (gg % p.tiles_b) * CBT + y;
        if (cs < a.Cs && cb < a.Cb) {
            float s = 0.f;
            for (int w = gg; w < CAE_NWARP; w += p.GP) s += red[w * NACC + rem];
            a.partials[(size_t)blockIdx.x * nelem + ((size_t)cs * a.Cb + cb) * KK + t] = s;
        }
    }
    if (cae_last_block(a.ticket)) {
        const int rows = gridDim.x;
        for (int e = tid; e < nelem; e += CAE_NT) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += __ldcg(a.partials + (size_t)r * nelem + e);
            a.grad[e] = s;
        }
    }
}

// =======================================================================================
// WGRAD v2b - "output parallel" (many channels, small spatial extent): a register-tiled GEMM
//   G[m][n] = sum_k A[k][m] * Bm[k][n],  m = cs, n = (cb, tap), k = position (n_s, i, j)
// A (small operand) and the im2col of the big operand are staged per 32-position slab.
// CTA tile 64 x 128, thread tile 4 x 8; blockIdx.x = position chunk (split-K, partial rows summed by
// the last CTA), blockIdx.y = (m tile, n tile).
// =======================================================================================
#define WG_BM 64
#define WG_BN 128
#define WG_KS 32
struct WgGemmPlan {
    int n_mtiles, n_ntiles;
    int kchunk;     // positions per blockIdx.x
    int nchunks;
    int KK, KW;
};

static __global__ void __launch_bounds__(CAE_NT) k_wgrad2b(const WgradArgs a, const WgGemmPlan p) {
    __shared__ __align__(16) float As[WG_KS][WG_BM + 4];
    __shared__ __align__(16) float Bs[WG_KS][WG_BN + 4];
    const int tid = threadIdx.x;
    const int mt = blockIdx.y / p.n_ntiles, nt = blockIdx.y - mt * p.n_ntiles;
    const int m0 = mt * WG_BM, n0 = nt * WG_BN;
    const int tm = tid >> 4, tn = tid & 15;      // 16 x 16 threads: rows tm*4.., cols tn*4.. and 64+tn*4..
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    const int Hs = sv.H, Ws = sv.W, HW = Hs * Ws;
    const int Ntot = a.Cb * p.KK;
    const int k_begin = blockIdx.x * p.kchunk;
    const int k_end = min(a.total, k_begin + p.kchunk);

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += WG_KS) {
        __syncthreads();
        // A slab: As[k][m] = S(pos k0+k, cs m0+m); lanes along k (consecutive positions are contiguous in memory)
        for (int i = tid; i < WG_KS * WG_BM; i += CAE_NT) {
            const int kk = i & (WG_KS - 1), m = i >> 5;
            const int pos = k0 + kk, cs = m0 + m;
            float v = 0.f;
            if (pos < k_end && cs < a.Cs) {
                const int n = pos / HW, r = pos - n * HW;
                const int ii = r / Ws, jj = r - ii * Ws;
                const ChanCoef kc = load_coef(a.sm, cs);
                v = src_value(a.sm, sbase + (long long)n * sv.sN + (long long)cs * sv.sC + (long long)ii * sv.ld + jj, kc);
            }
            As[kk][m] = v;
        }
        // B slab (im2col): Bs[k][nn] = B(pos -> (cb, ky, kx))
        for (int i = tid; i < WG_KS * WG_BN; i += CAE_NT) {
            const int kk = i & (WG_KS - 1), nn = i >> 5;
            const int pos = k0 + kk, col = n0 + nn;
            float v = 0.f;
            if (pos < k_end && col < Ntot) {
                const int cb = col / p.KK, t = col - cb * p.KK;
                const int ky = t / p.KW, kx = t - ky * p.KW;
                const int n = pos / HW, r = pos - n * HW;
                const int ii = r / Ws, jj = r - ii * Ws;
                const int y = ii * a.s + ky - a.p, x = jj * a.s + kx - a.p;
                if (y >= 0 && y < bv.H && x >= 0 && x < bv.W) {
                    const ChanCoef kc = load_coef(a.bg, cb);
                    v = src_value(a.bg, bbase + (long long)n * bv.sN + (long long)cb * bv.sC + (long long)y * bv.ld + x, kc);
                }
            }
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < WG_KS; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tn * 4]);
            const float am[4] = {av.x, av.y, av.z, av.w};
            const float bn[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
        }
    }
    const size_t nelem = (size_t)a.Cs * Ntot;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + (j < 4 ? tn * 4 + j : 64 + tn * 4 + (j - 4));
            if (m < a.Cs && col < Ntot) a.partials[(size_t)blockIdx.x * nelem + (size_t)m * Ntot + col] = acc[i][j];
        }
    }
    if (cae_last_block(a.ticket)) {
        const int rows = gridDim.x;
        for (size_t e = tid; e < nelem; e += CAE_NT) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += __ldcg(a.partials + (size_t)r * nelem + e);
            a.grad[e] = s;
        }
    }
}
