// ConvTranspose2d forward / input gradient / weight gradient as tensor-core GEMMs (tc_gemm.cu) for the layers whose
// channel counts make them a genuine dense contraction (BASELINE configs[3]: 1024->512, 512->256, 256->128, 128->64,
// 64->32 of the 4x64x64 -> 4x1024x1024 decoder; reference decoder.py:44-48 -> aten::convolution(transposed) and
// convolution_backward).  Activations stay fp32 NCHW outside (same CaeSrc / CaeView / CaeEpilogue contract as
// cae_conv_up / cae_conv_down / cae_conv_wgrad, so BatchNorm statistics, ReLU masks and on-load transforms are the
// shared epilogue code); inside, a layer is
//
//   forward   A[m, ci]  = transform(x)[n, ci, iy, ix]            m = (n, iy, ix)             k_tc_pack_act  (hi/lo split)
//             Wf[(t,co), ci] = W[ci, co, t]                       t = ky*kw + kx              k_tc_pack_wf
//             cols[m, (t,co)] = A Wf^T                            tcgen05 GEMM, K = Cin
//             y[n, co, oy, ox] = bias + sum over the taps that hit (oy, ox) of cols           k_tc_col2im + epilogue
//   backward  dcols[m, (t,co)] = transform(dy)[n, co, s*iy+ky, s*ix+kx]                       k_tc_im2col    (hi/lo split)
//             dx[m, ci] = dcols Wd^T,  Wd[ci, (t,co)] = W[ci, co, t]                          tcgen05 GEMM, K = T*Cout
//             dx -> NCHW + ReLU mask + BatchNorm-backward sums                                k_tc_unpack + epilogue
//             dW[ci, (t,co)] = sum_m A[m, ci] dcols[m, (t,co)]                                tcgen05 GEMM, MN-major operands,
//                                                                                             split-K, fixed-order reduce
// All reductions are fixed-order (split-K slices summed in index order; statistics through the two-stage partial rows
// of the conv family): results are bitwise reproducible.
#include "capi_host.h"
#include "conv_family.cuh"

extern "C" int cae_tc_gemm(const CaeTcGemm* g, void* stream);

namespace {

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    tf32_split(v, hi, lo);
}

// ---- forward operand: NCHW source (with its on-load transform) -> [m, ci] hi / lo ------------------------------------
__global__ void __launch_bounds__(256) k_tc_pack_act(const CaeSrc in, float* __restrict__ hi, float* __restrict__ lo,
                                                      long long lda, int P, int HW, int W) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const CaeView& iv = in.t0;
    const int C = iv.C;
    const int p = blockIdx.x * 32 + tx;
    const int c0 = blockIdx.y * 32;
    long long base = 0;
    if (p < P) {
        const int n = p / HW, r = p - n * HW, iy = r / W, ix = r - iy * W;
        base = src_cursor_offset(in) + (long long)n * iv.sN + (long long)iy * iv.ld + ix;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i;
        float v = 0.f;
        if (p < P && c < C) v = src_value(in, base + (long long)c * iv.sC, load_coef(in, c));
        tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pl = ty + 8 * i;
        const long long pp = (long long)blockIdx.x * 32 + pl;
        const int c = c0 + tx;
        if (pp < P && c < lda) {
            float h, l;
            split_tf32(c < C ? tile[tx][pl] : 0.f, h, l);
            hi[pp * lda + c] = h;
            lo[pp * lda + c] = l;
        }
    }
}

// ---- weights W[ci][co][t] -> forward operand Wf[(t,co)][ci] (transpose through shared memory) -----------------------
__global__ void __launch_bounds__(256) k_tc_pack_wf(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                                     int Cin, int Cout, int T, long long ldk) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Q = Cout * T;
    const int q0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty + 8 * i, q = q0 + tx;
        tile[ty + 8 * i][tx] = (ci < Cin && q < Q) ? __ldg(w + (size_t)ci * Q + q) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty + 8 * i, ci = ci0 + tx;
        if (q < Q && ci < ldk) {
            const int co = q / T, t = q - co * T;
            float h, l;
            split_tf32(ci < Cin ? tile[tx][ty + 8 * i] : 0.f, h, l);
            const size_t o = (size_t)(t * Cout + co) * ldk + ci;
            hi[o] = h;
            lo[o] = l;
        }
    }
}

// ---- weights W[ci][co][t] -> input-gradient operand Wd[ci][(t,co)] ----------------------------------------------------------
__global__ void __launch_bounds__(256) k_tc_pack_wd(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                                     int Cin, int Cout, int T, long long ldn) {
    const long long total = (long long)Cin * ldn;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i / ldn), col = (int)(i - (long long)ci * ldn);
        float v = 0.f;
        if (col < T * Cout) {
            const int t = col / Cout, co = col - t * Cout;
            v = __ldg(w + ((size_t)ci * Cout + co) * T + t);
        }
        float h, l;
        split_tf32(v, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// ---- forward tail: y[n, co, oy, ox] = bias + sum_{taps hitting (oy, ox)} cols[(n, iy, ix), (t, co)] + epilogue -------------
template <int COT>
__global__ void __launch_bounds__(CAE_NT) k_tc_col2im(const float* __restrict__ cols, long long ldn, const ConvArgs a, int Hin,
                                                       int Win) {
    const int co0 = blockIdx.y * COT;
    const int Hout = a.out.H, Wout = a.out.W;
    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;
    const bool vec = (COT % 4 == 0) && (a.Cout % 4 == 0) && (ldn % 4 == 0) && (co0 + COT <= a.Cout);
    for (int g = blockIdx.x * CAE_NT + threadIdx.x; g < a.total; g += gridDim.x * CAE_NT) {
        const int n = g / (Hout * Wout);
        const int r = g - n * (Hout * Wout);
        const int oy = r / Wout, ox = r - oy * Wout;
        float acc[COT];
#pragma unroll
        for (int j = 0; j < COT; ++j) acc[j] = 0.f;
        for (int ky = oy % a.s; ky < a.kh; ky += a.s) {
            const int iy = (oy - ky) / a.s;
            if (oy < ky || iy >= Hin) continue;
            for (int kx = ox % a.s; kx < a.kw; kx += a.s) {
                const int ix = (ox - kx) / a.s;
                if (ox < kx || ix >= Win) continue;
                const float* src = cols + ((long long)(n * Hin + iy) * Win + ix) * ldn + (long long)(ky * a.kw + kx) * a.Cout + co0;
                if (vec) {
#pragma unroll
                    for (int j = 0; j < COT; j += 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src + j));
                        acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < a.Cout) acc[j] += __ldg(src + j);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < COT; ++j)
            if (co0 + j < a.Cout)
                epi_element(a.epi, a.out, ech[j], n, co0 + j, oy, ox, acc[j], 0ll, a.inv_count, s1[j], s2[j]);
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}

// ---- the same tail for wide planes (Wout >= 48): a CTA walks tiles of one output row x 64 pixels x 32 channels.  Gather with
// lane = channel (every tap of a pixel is ONE coalesced 128-byte line of cols; the thread-per-pixel kernel above fetches
// 32-byte pieces scattered ldn floats apart), transpose through shared memory, epilogue + store with 8 consecutive pixels
// per channel and warp instruction.  Same tap order per pixel as above, so y is bit-identical; the BatchNorm sums are
// partitioned differently (8 threads per channel, fixed-order butterfly, one partial row per CTA).
__global__ void __launch_bounds__(CAE_NT) k_tc_col2im_tile(const float* __restrict__ cols, long long ldn, const ConvArgs a, int Hin,
                                                            int Win, int xtiles, int ntiles) {
    __shared__ float so[32][65];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cm = threadIdx.x >> 3, xl = threadIdx.x & 7;          // epilogue role: channel co0 + cm, pixels xl + 8 i
    const int co0 = blockIdx.y * 32;
    const int Hout = a.out.H, Wout = a.out.W, Cout = a.Cout;
    const bool cvalid = co0 + cm < Cout;
    const EpiCh ech = epi_load_channel(a.epi, co0 + cm, cvalid);
    float s1 = 0.f, s2 = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int xt = tile % xtiles, row = tile / xtiles;
        const int oy = row % Hout, n = row / Hout;
        const int ox0 = xt * 64, npx = min(64, Wout - ox0);
        // a warp owns pixels warp + 8 u: all <= 2 x 2 taps of all eight pixels are requested before the first add (32
        // independent 128-byte lines in flight per warp; with the taps fetched one after the other behind their bounds
        // checks a tile cost 23 us of exposed latency).  Missing taps contribute +0.f: same sums as the pixel kernel.
        float v[8][4];
        const bool cok = co0 + lane < Cout;
        const int ky0 = oy % a.s, iy0 = oy / a.s;                    // taps ky0 + s jy hit input rows iy0 - jy
        const float* cbase = cols + (long long)n * Hin * Win * ldn + co0 + lane;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ox = ox0 + warp + 8 * u;
            const int kx0 = ox % a.s, ix0 = ox / a.s;
#pragma unroll
            for (int jy = 0; jy < 2; ++jy)
#pragma unroll
                for (int jx = 0; jx < 2; ++jx) {
                    const int ky = ky0 + a.s * jy, kx = kx0 + a.s * jx, iy = iy0 - jy, ix = ix0 - jx;
                    const bool ok = cok && warp + 8 * u < npx && ky < a.kh && kx < a.kw && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
                    v[u][jy * 2 + jx] = ok ? __ldg(cbase + (long long)(iy * Win + ix) * ldn + (long long)(ky * a.kw + kx) * Cout) : 0.f;
                }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (warp + 8 * u < npx) so[lane][warp + 8 * u] = ((v[u][0] + v[u][1]) + v[u][2]) + v[u][3];
        __syncthreads();
        if (cvalid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int px = xl + 8 * i;
                if (px < npx) epi_element(a.epi, a.out, ech, n, co0 + cm, oy, ox0 + px, so[cm][px], 0ll, a.inv_count, s1, s2);
            }
        }
        __syncthreads();
    }
    if (epi_reduces(a.epi.mode)) {
        double d1 = (double)s1, d2 = (double)s2;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        if (xl == 0 && cvalid) {
            double* dst = a.epi.partials + ((size_t)blockIdx.x * Cout + co0 + cm) * 2;
            dst[0] = d1;
            dst[1] = d2;
        }
        if (cae_last_block(a.epi.ticket)) {
            const double count = (double)a.out.N * Hout * Wout;
            if (a.epi.mode == CAE_EPI_STATS) finalize_bn_forward(a.epi.bn, a.epi.partials, gridDim.x, count);
            else if (a.epi.mode == CAE_EPI_MASKSTATS) finalize_bn_backward(a.epi.bn, a.epi.partials, gridDim.x, count);
        }
    }
}

// ---- backward operand: dcols[m, (t,co)] = transform(dy)[n, co, s*iy + ky, s*ix + kx] -> hi / lo ------------------------
template <int COT>
__global__ void __launch_bounds__(CAE_NT) k_tc_im2col(const CaeSrc dy, float* __restrict__ hi, float* __restrict__ lo,
                                                       long long ldn, int kh, int kw, int s, int Hin, int Win, int M) {
    const CaeView& dv = dy.t0;
    const int Cout = dv.C;
    const int co0 = blockIdx.y * COT;
    ChanCoef kc[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) kc[j] = load_coef(dy, min(co0 + j, Cout - 1));
    const long long cur = src_cursor_offset(dy);
    const bool vec = (COT % 4 == 0) && (Cout % 4 == 0) && (ldn % 4 == 0) && (co0 + COT <= Cout);
    for (int m = blockIdx.x * CAE_NT + threadIdx.x; m < M; m += gridDim.x * CAE_NT) {
        const int n = m / (Hin * Win);
        const int r = m - n * (Hin * Win);
        const int iy = r / Win, ix = r - iy * Win;
        const long long base = cur + (long long)n * dv.sN + (long long)(iy * s) * dv.ld + ix * s;
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx) {
                float h[COT], l[COT];
#pragma unroll
                for (int j = 0; j < COT; ++j) {
                    float v = 0.f;
                    if (co0 + j < Cout) v = src_value(dy, base + (long long)(co0 + j) * dv.sC + (long long)ky * dv.ld + kx, kc[j]);
                    split_tf32(v, h[j], l[j]);
                }
                const long long o = (long long)m * ldn + (long long)(ky * kw + kx) * Cout + co0;
                if (vec) {
#pragma unroll
                    for (int j = 0; j < COT; j += 4) {
                        *reinterpret_cast<float4*>(hi + o + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
                        *reinterpret_cast<float4*>(lo + o + j) = make_float4(l[j], l[j + 1], l[j + 2], l[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < Cout) {
                            hi[o + j] = h[j];
                            lo[o + j] = l[j];
                        }
                }
            }
    }
}

// ---- the same operand for the layers with wide planes (Win >= 24): a CTA owns 32 consecutive positions of one input row
// and 32 channels.  The dy rows they touch (kh rows x (31 s + kw) columns per channel) are read ONCE, coalesced, transformed
// and parked in shared memory; the (position, tap) rows of dcols then leave as full 128-byte lines (lane = channel).  The
// thread-per-position kernel above reads dy with stride s and scatters 32-byte pieces ldn floats apart: 1.5 TB/s of
// traffic at the 64 -> 32 layer of BASELINE configs[3] (1.12 ms); this one is bit-identical and HBM-bound.
__global__ void __launch_bounds__(CAE_NT) k_tc_im2col_tile(const CaeSrc dy, float* __restrict__ hi, float* __restrict__ lo,
                                                            long long ldn, int kh, int kw, int s, int Hin, int Win, int xtiles,
                                                            int WT, int cpitch) {
    extern __shared__ float sv[];                               // [32 channels][cpitch >= kh * WT, odd]
    const CaeView& dv = dy.t0;
    const int Cout = dv.C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.y * 32;
    int b = blockIdx.x;
    const int xt = b % xtiles;
    b /= xtiles;
    const int iy = b % Hin, n = b / Hin;
    const int ix0 = xt * 32;
    const int npos = min(32, Win - ix0);
    const int wt = (npos - 1) * s + kw;                         // columns of dy this tile touches
    const long long base = src_cursor_offset(dy) + (long long)n * dv.sN + (long long)(iy * s) * dv.ld + ix0 * s;
    // rows (channel, ky): a warp per row, lanes over the columns
    for (int r = warp; r < 32 * kh; r += CAE_NWARP) {
        const int c = r / kh, ky = r - c * kh;
        if (co0 + c >= Cout) continue;
        const ChanCoef kc = load_coef(dy, co0 + c);
        const long long roff = base + (long long)(co0 + c) * dv.sC + (long long)ky * dv.ld;
        float* dst = sv + c * cpitch + ky * WT;
        for (int x = lane; x < wt; x += 32) dst[x] = src_value(dy, roff + x, kc);
    }
    __syncthreads();
    if (co0 + lane >= Cout) return;
    const int T = kh * kw;
    const long long m0 = ((long long)n * Hin + iy) * Win + ix0;
    const float* src = sv + lane * cpitch;
    for (int e = warp; e < npos * T; e += CAE_NWARP) {
        const int pos = e / T, t = e - pos * T;
        const int ky = t / kw, kx = t - ky * kw;
        float h, l;
        split_tf32(src[ky * WT + pos * s + kx], h, l);
        const long long o = (m0 + pos) * ldn + (long long)t * Cout + co0 + lane;
        hi[o] = h;
        lo[o] = l;
    }
}

// ---- input-gradient tail: dx[m, ci] (GEMM output) -> NCHW + epilogue (ReLU mask, BatchNorm-backward sums) ----------------
__global__ void __launch_bounds__(256) k_tc_unpack(const float* __restrict__ dxg, long long ldx, const ConvArgs a, int P) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 1-D block: the reduction helpers index by threadIdx.x
    const int C = a.out.C, H = a.out.H, W = a.out.W;
    const int c0 = blockIdx.y * 32;
    EpiCh ech[4];
    float s1[4], s2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty * 4 + i;
        ech[i] = epi_load_channel(a.epi, c, c < C);
        s1[i] = s2[i] = 0.f;
    }
    for (int p0 = blockIdx.x * 32; p0 < P; p0 += gridDim.x * 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pl = ty + 8 * i;
            const int c = c0 + tx;
            tile[pl][tx] = (p0 + pl < P && c < C) ? __ldg(dxg + (long long)(p0 + pl) * ldx + c) : 0.f;
        }
        __syncthreads();
        const int p = p0 + tx;
        if (p < P) {
            const int n = p / (H * W), r = p - n * (H * W), iy = r / W, ix = r - iy * W;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c0 + ty * 4 + i;
                if (c < C) epi_element(a.epi, a.out, ech[i], n, c, iy, ix, tile[tx][ty * 4 + i], 0ll, a.inv_count, s1[i], s2[i]);
            }
        }
        __syncthreads();
    }
    if (epi_reduces(a.epi.mode)) {
        // warp ty owns channels c0 + 4*ty .. +3: one fixed-order butterfly per channel, lane 0 writes the partial row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double d1 = warp_sum_d((double)s1[i]), d2 = warp_sum_d((double)s2[i]);
            const int c = c0 + ty * 4 + i;
            if (tx == 0 && c < C) {
                double* dst = a.epi.partials + ((size_t)blockIdx.x * C + c) * 2;
                dst[0] = d1;
                dst[1] = d2;
            }
        }
        if (cae_last_block(a.epi.ticket)) {
            const double count = (double)a.out.N * H * W;
            if (a.epi.mode == CAE_EPI_STATS) finalize_bn_forward(a.epi.bn, a.epi.partials, gridDim.x, count);
            else if (a.epi.mode == CAE_EPI_MASKSTATS) finalize_bn_backward(a.epi.bn, a.epi.partials, gridDim.x, count);
        }
    }
}

// ---- weight-gradient tail: dW[ci][co][t] = sum over split-K slices (index order) of part[s][ci][(t,co)] ----------------------
__global__ void __launch_bounds__(256) k_tc_wgrad_reduce(const float* __restrict__ part, long long split_stride, int splits,
                                                          long long ldn, float* __restrict__ grad, int Cin, int Cout, int T) {
    const long long total = (long long)Cin * Cout * T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // i enumerates (ci, t, co) so that the reads are contiguous; the write goes to the native (ci, co, t) slot
        const int ci = (int)(i / ((long long)Cout * T));
        const int rem = (int)(i - (long long)ci * Cout * T);
        const int t = rem / Cout, co = rem - t * Cout;
        const float* src = part + (long long)ci * ldn + rem;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(src + (long long)k * split_stride);
        grad[((size_t)ci * Cout + co) * T + t] = s;
    }
}

int pick_tile_n(int N) {
    if (N >= 512 && ((N + 255) / 256 * 256 - N) < 128) return 256;
    return 128;
}

int gemm(int M, int N, int K, const float* ah, const float* al, long long lda, int a_mn, const float* bh, const float* bl,
         long long ldb, int b_mn, float* C, long long ldc, int splits, long long split_stride, void* stream) {
    CaeTcGemm g{};
    g.M = M; g.N = N; g.K = K;
    g.a_hi = ah; g.a_lo = al; g.lda = lda; g.a_mn_major = a_mn;
    g.b_hi = bh; g.b_lo = bl; g.ldb = ldb; g.b_mn_major = b_mn;
    g.C = C; g.ldc = ldc; g.splits = splits; g.split_stride = split_stride;
    g.tile_n = pick_tile_n(N);
    return cae_tc_gemm(&g, stream);
}

int check_desc(const CaeTcConv* c) {
    CAE_REQUIRE(c, "tc_conv: null descriptor");
    CAE_REQUIRE(c->Cin > 0 && c->Cout > 0 && c->kh > 0 && c->kw > 0 && c->stride > 0 && c->N > 0 && c->Hin > 0 && c->Win > 0,
                "tc_conv: bad geometry");
    CAE_REQUIRE(c->Hout >= (c->Hin - 1) * c->stride + c->kh && c->Wout >= (c->Win - 1) * c->stride + c->kw,
                "tc_conv: output %dx%d smaller than (in-1)*stride + kernel", c->Hout, c->Wout);
    CAE_REQUIRE(c->lda >= c->Cin && c->lda % 4 == 0 && c->ldn >= (long long)c->kh * c->kw * c->Cout && c->ldn % 4 == 0,
                "tc_conv: lda / ldn must cover Cin / kh*kw*Cout and be multiples of 4");
    CAE_REQUIRE(c->a_hi && c->a_lo && c->w_hi && c->w_lo && c->cols, "tc_conv: workspace pointers missing");
    CAE_REQUIRE((long long)c->N * c->Hin * c->Win < (1ll << 31) && (long long)c->N * c->Cout * c->Hout * c->Wout < (1ll << 31),
                "tc_conv: tensor too large for 32-bit positions");
    return CAE_OK;
}

}  // namespace

extern "C" int cae_tc_convt_supported(int Cin, int Cout, int kh, int kw, int stride, int pad) {
    return (pad == 0 && Cin >= 32 && Cin % 4 == 0 && Cout % 4 == 0 && Cout >= 8 && kh >= 1 && kw >= 1 && stride >= 1) ? 1 : 0;
}

extern "C" long long cae_tc_convt_wgrad_splits(const CaeTcConv* c) {
    if (!c) return -1;
    const int TC = c->kh * c->kw * c->Cout;
    const long long P = (long long)c->N * c->Hin * c->Win;
    const int bn = pick_tile_n(TC);
    const long long tiles = (long long)((c->Cin + 127) / 128) * ((TC + bn - 1) / bn);
    const long long nkb = (P + 31) / 32;
    long long splits = (2 * CAE_NUM_SMS + tiles - 1) / tiles;
    // at most 64 K blocks (2048 positions) per slice.  Introduced when one TMEM accumulator ran over a slice's whole K range
    // (truncation error linear in K); the GEMM now promotes every 2 K blocks into round-to-nearest register sums, so the
    // cap is no longer needed for accuracy - it is kept because the extra slices cost < 20 MB of partial tiles and the
    // parity fixtures were recorded with it
    if (splits < (nkb + 63) / 64) splits = (nkb + 63) / 64;
    if (splits > nkb / 8) splits = nkb / 8;
    if (splits < 1) splits = 1;
    const long long per = (nkb + splits - 1) / splits;
    return (nkb + per - 1) / per;
}

extern "C" int cae_tc_convt_fwd(const CaeTcConv* c, const CaeSrc* in, const float* weight, const CaeView* out,
                                const CaeEpilogue* epi, void* stream) {
    int rc;
    if ((rc = check_desc(c))) return rc;
    CAE_REQUIRE(in && weight && out && epi, "tc_convT_fwd: null argument");
    CAE_REQUIRE(in->t0.N == c->N && in->t0.C == c->Cin && in->t0.H == c->Hin && in->t0.W == c->Win && out->N == c->N &&
                    out->C == c->Cout && out->H == c->Hout && out->W == c->Wout,
                "tc_convT_fwd: views do not match the descriptor");
    CAE_REQUIRE(epi->mode == CAE_EPI_PLAIN || epi->mode == CAE_EPI_STATS, "tc_convT_fwd: epilogue mode %d not supported", epi->mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int T = c->kh * c->kw, TC = T * c->Cout;
    const int P = c->N * c->Hin * c->Win;
    CAE_REQUIRE(c->cols_len >= (long long)P * c->ldn, "tc_convT_fwd: cols scratch too small");
    k_tc_pack_act<<<dim3((P + 31) / 32, (c->lda + 31) / 32), dim3(32, 8), 0, st>>>(*in, c->a_hi, c->a_lo, c->lda, P,
                                                                                 c->Hin * c->Win, c->Win);
    if ((rc = cae_check_launch("k_tc_pack_act"))) return rc;
    k_tc_pack_wf<<<dim3((c->Cout * T + 31) / 32, (c->lda + 31) / 32), dim3(32, 8), 0, st>>>(weight, c->w_hi, c->w_lo, c->Cin,
                                                                                            c->Cout, T, c->lda);
    if ((rc = cae_check_launch("k_tc_pack_wf"))) return rc;
    if ((rc = gemm(P, TC, c->Cin, c->a_hi, c->a_lo, c->lda, 0, c->w_hi, c->w_lo, c->lda, 0, c->cols, c->ldn, 1, 0, stream)))
        return rc;
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.kh = c->kh; a.kw = c->kw; a.s = c->stride; a.p = 0;
    a.out = *out;
    a.epi = *epi;
    a.Cin = c->Cin; a.Cout = c->Cout;
    a.total = c->N * c->Hout * c->Wout;
    a.inv_count = 1.f;
    if (epi_reduces(a.epi.mode)) CAE_REQUIRE(a.epi.partials && a.epi.ticket, "tc_convT_fwd: reducing epilogue needs partials + ticket");
    const long long ntiles = (long long)c->N * c->Hout * ceil_div(c->Wout, 64);
    if (c->Wout >= 48 && ntiles < (1ll << 31) && c->kh <= 2 * c->stride && c->kw <= 2 * c->stride) {
        dim3 grid((unsigned)min(ntiles, (long long)CAE_MAX_GRID_X), ceil_div(c->Cout, 32));
        k_tc_col2im_tile<<<grid, CAE_NT, 0, st>>>(c->cols, c->ldn, a, c->Hin, c->Win, ceil_div(c->Wout, 64), (int)ntiles);
        return cae_check_launch("k_tc_col2im_tile");
    }
    dim3 grid(min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X), ceil_div(c->Cout, 8));
    k_tc_col2im<8><<<grid, CAE_NT, 0, st>>>(c->cols, c->ldn, a, c->Hin, c->Win);
    return cae_check_launch("k_tc_col2im");
}

extern "C" int cae_tc_convt_im2col(const CaeTcConv* c, const CaeSrc* dy, void* stream) {
    int rc;
    if ((rc = check_desc(c))) return rc;
    CAE_REQUIRE(dy && c->dcols_hi && c->dcols_lo, "tc_convT_im2col: null argument");
    CAE_REQUIRE(dy->t0.N == c->N && dy->t0.C == c->Cout && dy->t0.H == c->Hout && dy->t0.W == c->Wout,
                "tc_convT_im2col: dy view does not match the descriptor");
    const int M = c->N * c->Hin * c->Win;
    const int WT = 31 * c->stride + c->kw, cpitch = (c->kh * WT) | 1;
    const size_t smem = (size_t)32 * cpitch * sizeof(float);
    const long long tiles = (long long)c->N * c->Hin * ceil_div(c->Win, 32);
    if (c->Win >= 24 && smem <= 96 * 1024 && tiles < (1ll << 31)) {
        ensure_smem_limit(k_tc_im2col_tile, 96 * 1024);
        dim3 grid((unsigned)tiles, ceil_div(c->Cout, 32));
        k_tc_im2col_tile<<<grid, CAE_NT, smem, (cudaStream_t)stream>>>(*dy, c->dcols_hi, c->dcols_lo, c->ldn, c->kh, c->kw, c->stride,
                                                                     c->Hin, c->Win, ceil_div(c->Win, 32), WT, cpitch);
        return cae_check_launch("k_tc_im2col_tile");
    }
    dim3 grid(min(ceil_div(M, CAE_NT), CAE_MAX_GRID_X * 4), ceil_div(c->Cout, 8));
    k_tc_im2col<8><<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(*dy, c->dcols_hi, c->dcols_lo, c->ldn, c->kh, c->kw, c->stride,
                                                             c->Hin, c->Win, M);
    return cae_check_launch("k_tc_im2col");
}

extern "C" int cae_tc_convt_dgrad(const CaeTcConv* c, const float* weight, const CaeView* dx, const CaeEpilogue* epi,
                                  void* stream) {
    int rc;
    if ((rc = check_desc(c))) return rc;
    CAE_REQUIRE(weight && dx && epi && c->dcols_hi && c->dcols_lo, "tc_convT_dgrad: null argument");
    CAE_REQUIRE(dx->N == c->N && dx->C == c->Cin && dx->H == c->Hin && dx->W == c->Win, "tc_convT_dgrad: dx view mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = c->kh * c->kw, TC = T * c->Cout;
    const int P = c->N * c->Hin * c->Win;
    CAE_REQUIRE(c->cols_len >= (long long)P * c->lda, "tc_convT_dgrad: cols scratch too small");
    k_tc_pack_wd<<<min(ceil_div((long long)c->Cin * c->ldn, 256), CAE_NUM_SMS * 8), 256, 0, st>>>(weight, c->w_hi, c->w_lo, c->Cin,
                                                                                                  c->Cout, T, c->ldn);
    if ((rc = cae_check_launch("k_tc_pack_wd"))) return rc;
    if ((rc = gemm(P, c->Cin, TC, c->dcols_hi, c->dcols_lo, c->ldn, 0, c->w_hi, c->w_lo, c->ldn, 0, c->cols, c->lda, 1, 0, stream)))
        return rc;
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.out = *dx;
    a.epi = *epi;
    if (a.epi.mode == CAE_EPI_MASKSTATS && a.epi.act.p == nullptr) a.epi.mode = CAE_EPI_PLAIN;
    CAE_REQUIRE(a.epi.mode == CAE_EPI_PLAIN || a.epi.mode == CAE_EPI_MASK || a.epi.mode == CAE_EPI_MASKSTATS,
                "tc_convT_dgrad: epilogue mode %d not supported", a.epi.mode);
    if (epi_reduces(a.epi.mode)) CAE_REQUIRE(a.epi.partials && a.epi.ticket, "tc_convT_dgrad: reducing epilogue needs partials + ticket");
    a.Cin = a.Cout = c->Cin;
    a.total = P;
    a.inv_count = 1.f;
    dim3 grid(min((P + 31) / 32, CAE_MAX_GRID_X), (c->Cin + 31) / 32);
    k_tc_unpack<<<grid, 256, 0, st>>>(c->cols, c->lda, a, P);
    return cae_check_launch("k_tc_unpack");
}

extern "C" int cae_tc_convt_wgrad(const CaeTcConv* c, float* grad, void* stream) {
    int rc;
    if ((rc = check_desc(c))) return rc;
    CAE_REQUIRE(grad && c->dcols_hi && c->dcols_lo, "tc_convT_wgrad: null argument");
    const int T = c->kh * c->kw, TC = T * c->Cout;
    const int P = c->N * c->Hin * c->Win;
    const int splits = (int)cae_tc_convt_wgrad_splits(c);
    const long long split_stride = (long long)c->Cin * c->ldn;
    CAE_REQUIRE(c->cols_len >= split_stride * splits, "tc_convT_wgrad: cols scratch too small (%d splits)", splits);
    if ((rc = gemm(c->Cin, TC, P, c->a_hi, c->a_lo, c->lda, 1, c->dcols_hi, c->dcols_lo, c->ldn, 1, c->cols, c->ldn, splits,
                   split_stride, stream)))
        return rc;
    const long long total = (long long)c->Cin * TC;
    k_tc_wgrad_reduce<<<min(ceil_div(total, 256), CAE_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(c->cols, split_stride, splits,
                                                                                                    c->ldn, grad, c->Cin, c->Cout, T);
    return cae_check_launch("k_tc_wgrad_reduce");
}
