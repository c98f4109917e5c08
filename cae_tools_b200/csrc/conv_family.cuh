// Direct (SIMT fp32) convolution family: strided "down" conv, transposed "up" conv, weight gradient,
// all with on-load operand transforms (BatchNorm+ReLU apply / BatchNorm-backward affine) and fused
// epilogues (bias, BatchNorm statistics, ReLU-mask + BatchNorm-backward sums, sigmoid, MSE loss+grad).
#pragma once
#include "common.cuh"

struct ConvArgs {
    CaeSrc in;
    const float* w;
    int kh, kw, s, p;
    CaeView out;
    CaeEpilogue epi;
    int Cin, Cout;
    int QH, QW;      // up: cell grid (ceil((Hout+p)/s) etc.); down: Hout, Wout
    int total;       // positions = N * QH * QW
    int ci_chunk;    // input channels staged in shared memory per pass
    float inv_count; // 1 / (N*C*H*W) of the output (MSE)
};

// per output-channel constants of the epilogue
struct EpiCh {
    float bias, scale, shift, mean, invstd;
    ChanCoef tk;
    ChanCoef ak;   // addend coefficients
};

__device__ __forceinline__ EpiCh epi_load_channel(const CaeEpilogue& e, int co, bool ok) {
    EpiCh c;
    c.bias = (ok && e.bias) ? __ldg(e.bias + co) : 0.f;
    c.scale = c.shift = c.mean = c.invstd = 0.f;
    c.tk.k0 = 1.f; c.tk.k1 = 0.f; c.tk.k2 = 0.f;
    if (ok && e.mode == CAE_EPI_MASKSTATS) {
        c.scale = e.bn.scale[co];
        c.shift = e.bn.shift[co];
        c.mean = e.bn.mean[co];
        c.invstd = e.bn.invstd[co];
    }
    if (ok && e.mode == CAE_EPI_SIGMOID_MSE) c.tk = load_coef(e.target, co);
    c.ak.k0 = 1.f; c.ak.k1 = 0.f; c.ak.k2 = 0.f;
    if (ok && e.addend.t0.p) c.ak = load_coef(e.addend, co);
    return c;
}

// Apply the epilogue to one accumulated output element.
__device__ __forceinline__ void epi_element(const CaeEpilogue& e, const CaeView& out, const EpiCh& ch, int n, int co,
                                            int oy, int ox, float acc, long long tgt_base, float inv_count,
                                            float& s1, float& s2) {
    const long long off = (long long)n * out.sN + (long long)co * out.sC + (long long)oy * out.ld + ox;
    if (e.addend.t0.p) {
        const CaeView& av = e.addend.t0;
        acc += src_value(e.addend, (long long)n * av.sN + (long long)co * av.sC + (long long)oy * av.ld + ox, ch.ak);
    }
    switch (e.mode) {
        case CAE_EPI_PLAIN:
            out.p[off] = acc + ch.bias;
            break;
        case CAE_EPI_MASK: {
            const CaeView& a = e.act;
            float yp = __ldg(a.p + (long long)n * a.sN + (long long)co * a.sC + (long long)oy * a.ld + ox);
            out.p[off] = yp > 0.f ? acc : 0.f;
        } break;
        case CAE_EPI_STATS: {
            float v = acc + ch.bias;
            out.p[off] = v;
            s1 += v;
            s2 = fmaf(v, v, s2);
        } break;
        case CAE_EPI_MASKSTATS: {
            const CaeView& a = e.act;
            float yp = __ldg(a.p + (long long)n * a.sN + (long long)co * a.sC + (long long)oy * a.ld + ox);
            float z = fmaf(yp, ch.scale, ch.shift);
            float dz = z > 0.f ? acc : 0.f;
            out.p[off] = dz;
            float xh = (yp - ch.mean) * ch.invstd;
            s1 += dz;
            s2 = fmaf(dz, xh, s2);
        } break;
        case CAE_EPI_SIGMOID: {
            float v = acc + ch.bias;
            out.p[off] = 1.f / (1.f + expf(-v));
        } break;
        case CAE_EPI_SIGMOID_MSE: {
            float v = acc + ch.bias;
            float yh = 1.f / (1.f + expf(-v));
            const CaeView& t = e.target.t0;
            long long toff = tgt_base + (long long)n * t.sN + (long long)co * t.sC + (long long)oy * t.ld + ox;
            float d = yh - src_value(e.target, toff, ch.tk);
            s2 = fmaf(d, d, s2);
            float dz = 2.f * d * inv_count * yh * (1.f - yh);
            s1 += dz;
            if (e.write_mode == 0) out.p[off] = dz;
            else if (e.write_mode == 1) out.p[off] = yh;
        } break;
    }
}

__host__ __device__ __forceinline__ bool epi_reduces(int mode) {
    return mode == CAE_EPI_STATS || mode == CAE_EPI_MASKSTATS || mode == CAE_EPI_SIGMOID_MSE;
}

// loss = sum_c Q_c / count ; dbias[c] = S1_c
__device__ __forceinline__ void finalize_mse(const CaeEpilogue& e, const double* part, int rows, int C, double count) {
    __shared__ double qc[64];
    __shared__ double qtot;
    if (threadIdx.x == 0) qtot = 0.0;
    __syncthreads();
    // channels in passes of <= 64 so the per-channel losses can be added in channel order
    for (int cb = 0; cb < C; cb += 64) {
        const int cn = min(64, C - cb);
        for_each_channel_sums(part + (size_t)cb * 2, rows, C, [&](int c, double S1, double Q) {
            if (e.dbias) e.dbias[cb + c] = (float)S1;
            qc[c] = Q;
        }, cn);
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = qtot;
            for (int c = 0; c < cn; ++c) t += qc[c];
            qtot = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int slot = e.target.cursor ? __ldg(e.target.cursor) : 0;
        const double cs = e.count_scale > 0.f ? (double)e.count_scale : 1.0;
        if (e.loss_out) e.loss_out[slot] = (float)(qtot / count * cs);
    }
}

// tail shared by every kernel with a reducing epilogue: write this CTA's partial row, and let the
// last CTA of the grid finish the job.  COT channels per CTA starting at co0.
template <int COT>
__device__ __forceinline__ void epi_reduce_tail(const CaeEpilogue& e, const CaeView& out, int co0, float (&s1)[COT],
                                                float (&s2)[COT]) {
    const int C = out.C;
    float v[COT * 2];
#pragma unroll
    for (int j = 0; j < COT; ++j) {
        v[2 * j] = s1[j];
        v[2 * j + 1] = s2[j];
    }
    int nvalid = C - co0;
    nvalid = (nvalid > COT ? COT : nvalid) * 2;
    cta_reduce_store<COT * 2>(v, e.partials + ((size_t)blockIdx.x * C + co0) * 2, nvalid);
    if (cae_last_block(e.ticket)) {
        const double count = (double)out.N * out.H * out.W;
        if (e.mode == CAE_EPI_STATS) finalize_bn_forward(e.bn, e.partials, gridDim.x, count);
        else if (e.mode == CAE_EPI_MASKSTATS) finalize_bn_backward(e.bn, e.partials, gridDim.x, count);
        else finalize_mse(e, e.partials, gridDim.x, C, count * C);
    }
}

// =======================================================================================
// UP: transposed convolution, gather form.  One thread owns one "cell" (an S x S block of
// output pixels that share the same input neighbourhood) for COT output channels.
// weights [Cin][Cout][KH][KW]; staged in smem as [ci][tap][COT].
// =======================================================================================
template <int KH, int KW, int S, int COT>
__global__ void __launch_bounds__(CAE_NT) k_conv_up(const ConvArgs a) {
    constexpr int JY = (KH + S - 1) / S, JX = (KW + S - 1) / S, KK = KH * KW;
    extern __shared__ float sw[];
    const int co0 = blockIdx.y * COT;
    const int tid = threadIdx.x;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    const int Hin = iv.H, Win = iv.W, Hout = a.out.H, Wout = a.out.W;
    const bool single = a.ci_chunk >= a.Cin;

    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);

    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;

    auto stage_weights = [&](int c0, int cn) {
        for (int i = tid; i < cn * KK * COT; i += CAE_NT) {
            int j = i % COT, t = (i / COT) % KK, cl = i / (COT * KK);
            int co = co0 + j;
            sw[i] = co < a.Cout ? __ldg(a.w + ((size_t)(c0 + cl) * a.Cout + co) * KK + t) : 0.f;
        }
    };
    if (single) {
        stage_weights(0, a.Cin);
        __syncthreads();
    }

    for (int g0 = blockIdx.x * CAE_NT; g0 < a.total; g0 += gridDim.x * CAE_NT) {
        const int g = g0 + tid;
        const bool valid = g < a.total;
        int n = 0, qy = 0, qx = 0;
        if (valid) {
            n = g / (a.QH * a.QW);
            int r = g - n * (a.QH * a.QW);
            qy = r / a.QW;
            qx = r - qy * a.QW;
        }
        float acc[COT][S][S];
#pragma unroll
        for (int j = 0; j < COT; ++j)
#pragma unroll
            for (int py = 0; py < S; ++py)
#pragma unroll
                for (int px = 0; px < S; ++px) acc[j][py][px] = 0.f;

        // input neighbourhood offsets / validity (independent of ci)
        long long noff[JY][JX];
        bool nok[JY][JX];
#pragma unroll
        for (int jy = 0; jy < JY; ++jy)
#pragma unroll
            for (int jx = 0; jx < JX; ++jx) {
                int iy = qy - jy, ix = qx - jx;
                nok[jy][jx] = valid && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
                noff[jy][jx] = in_base + (long long)n * iv.sN + (long long)iy * iv.ld + ix;
            }

        for (int c0 = 0; c0 < a.Cin; c0 += a.ci_chunk) {
            const int cn = min(a.ci_chunk, a.Cin - c0);
            if (!single) {
                __syncthreads();
                stage_weights(c0, cn);
                __syncthreads();
            }
            for (int cl = 0; cl < cn; ++cl) {
                const int ci = c0 + cl;
                const ChanCoef kc = load_coef(a.in, ci);
                const long long coff = (long long)ci * iv.sC;
                float v[JY][JX];
#pragma unroll
                for (int jy = 0; jy < JY; ++jy)
#pragma unroll
                    for (int jx = 0; jx < JX; ++jx)
                        v[jy][jx] = nok[jy][jx] ? src_value(a.in, noff[jy][jx] + coff, kc) : 0.f;
                const float* wp = sw + cl * KK * COT;
#pragma unroll
                for (int py = 0; py < S; ++py)
#pragma unroll
                    for (int jy = 0; jy < JY; ++jy) {
                        const int ky = py + jy * S;
                        if (ky < KH) {
#pragma unroll
                            for (int px = 0; px < S; ++px)
#pragma unroll
                                for (int jx = 0; jx < JX; ++jx) {
                                    const int kx = px + jx * S;
                                    if (kx < KW) {
                                        const float* wt = wp + (ky * KW + kx) * COT;
#pragma unroll
                                        for (int j = 0; j < COT; ++j)
                                            acc[j][py][px] = fmaf(v[jy][jx], wt[j], acc[j][py][px]);
                                    }
                                }
                        }
                    }
            }
        }

        // epilogue
#pragma unroll
        for (int py = 0; py < S; ++py) {
            const int oy = qy * S + py - a.p;
#pragma unroll
            for (int px = 0; px < S; ++px) {
                const int ox = qx * S + px - a.p;
                if (valid && oy >= 0 && oy < Hout && ox >= 0 && ox < Wout) {
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < a.Cout)
                            epi_element(a.epi, a.out, ech[j], n, co0 + j, oy, ox, acc[j][py][px], tgt_base,
                                        a.inv_count, s1[j], s2[j]);
                }
            }
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}

// generic-geometry fallback (any kh, kw, stride, pad): one output pixel per thread, one channel per
// blockIdx.y, weights straight from global/L1.
static __global__ void __launch_bounds__(CAE_NT) k_conv_up_generic(const ConvArgs a) {
    const int co = blockIdx.y;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    const int Hout = a.out.H, Wout = a.out.W, KK = a.kh * a.kw;
    EpiCh ech = epi_load_channel(a.epi, co, true);
    float s1[1] = {0.f}, s2[1] = {0.f};
    for (int g0 = blockIdx.x * CAE_NT; g0 < a.total; g0 += gridDim.x * CAE_NT) {
        const int g = g0 + threadIdx.x;
        if (g < a.total) {
            int n = g / (Hout * Wout);
            int r = g - n * (Hout * Wout);
            int oy = r / Wout, ox = r - oy * Wout;
            const int ty = oy + a.p, tx = ox + a.p;
            float acc = 0.f;
            for (int ci = 0; ci < a.Cin; ++ci) {
                const ChanCoef kc = load_coef(a.in, ci);
                const float* wc = a.w + ((size_t)ci * a.Cout + co) * KK;
                for (int ky = ty % a.s; ky < a.kh; ky += a.s) {
                    int iy = (ty - ky) / a.s;
                    if (ty - ky < 0 || iy >= iv.H) continue;
                    for (int kx = tx % a.s; kx < a.kw; kx += a.s) {
                        int ix = (tx - kx) / a.s;
                        if (tx - kx < 0 || ix >= iv.W) continue;
                        long long off = in_base + (long long)n * iv.sN + (long long)ci * iv.sC + (long long)iy * iv.ld + ix;
                        acc = fmaf(src_value(a.in, off, kc), __ldg(wc + ky * a.kw + kx), acc);
                    }
                }
            }
            epi_element(a.epi, a.out, ech, n, co, oy, ox, acc, tgt_base, a.inv_count, s1[0], s2[0]);
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<1>(a.epi, a.out, co, s1, s2);
}

// =======================================================================================
// DOWN: strided convolution.  One thread owns one output pixel for COT output channels.
// weights [Cout][Cin][KH][KW]; staged in smem as [ci][tap][COT].
// =======================================================================================
template <int KH, int KW, int S, int COT>
__global__ void __launch_bounds__(CAE_NT) k_conv_down(const ConvArgs a) {
    constexpr int KK = KH * KW;
    extern __shared__ float sw[];
    const int co0 = blockIdx.y * COT;
    const int tid = threadIdx.x;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const int Hin = iv.H, Win = iv.W;
    const bool single = a.ci_chunk >= a.Cin;

    EpiCh ech[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) ech[j] = epi_load_channel(a.epi, co0 + j, co0 + j < a.Cout);
    float s1[COT], s2[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) s1[j] = s2[j] = 0.f;

    auto stage_weights = [&](int c0, int cn) {
        for (int i = tid; i < cn * KK * COT; i += CAE_NT) {
            int j = i % COT, t = (i / COT) % KK, cl = i / (COT * KK);
            int co = co0 + j;
            sw[i] = co < a.Cout ? __ldg(a.w + ((size_t)co * a.Cin + (c0 + cl)) * KK + t) : 0.f;
        }
    };
    if (single) {
        stage_weights(0, a.Cin);
        __syncthreads();
    }

    for (int g0 = blockIdx.x * CAE_NT; g0 < a.total; g0 += gridDim.x * CAE_NT) {
        const int g = g0 + tid;
        const bool valid = g < a.total;
        int n = 0, oy = 0, ox = 0;
        if (valid) {
            n = g / (a.QH * a.QW);
            int r = g - n * (a.QH * a.QW);
            oy = r / a.QW;
            ox = r - oy * a.QW;
        }
        float acc[COT];
#pragma unroll
        for (int j = 0; j < COT; ++j) acc[j] = 0.f;

        const int iy0 = oy * S - a.p, ix0 = ox * S - a.p;
        bool rok[KH], cok[KW];
#pragma unroll
        for (int ky = 0; ky < KH; ++ky) rok[ky] = valid && (iy0 + ky) >= 0 && (iy0 + ky) < Hin;
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) cok[kx] = (ix0 + kx) >= 0 && (ix0 + kx) < Win;
        const long long base = in_base + (long long)n * iv.sN + (long long)iy0 * iv.ld + ix0;

        for (int c0 = 0; c0 < a.Cin; c0 += a.ci_chunk) {
            const int cn = min(a.ci_chunk, a.Cin - c0);
            if (!single) {
                __syncthreads();
                stage_weights(c0, cn);
                __syncthreads();
            }
            for (int cl = 0; cl < cn; ++cl) {
                const int ci = c0 + cl;
                const ChanCoef kc = load_coef(a.in, ci);
                const long long coff = base + (long long)ci * iv.sC;
                const float* wp = sw + cl * KK * COT;
#pragma unroll
                for (int ky = 0; ky < KH; ++ky)
#pragma unroll
                    for (int kx = 0; kx < KW; ++kx) {
                        float v = (rok[ky] && cok[kx]) ? src_value(a.in, coff + (long long)ky * iv.ld + kx, kc) : 0.f;
                        const float* wt = wp + (ky * KW + kx) * COT;
#pragma unroll
                        for (int j = 0; j < COT; ++j) acc[j] = fmaf(v, wt[j], acc[j]);
                    }
            }
        }
        if (valid) {
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < a.Cout)
                    epi_element(a.epi, a.out, ech[j], n, co0 + j, oy, ox, acc[j], 0ll, a.inv_count, s1[j], s2[j]);
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<COT>(a.epi, a.out, co0, s1, s2);
}

static __global__ void __launch_bounds__(CAE_NT) k_conv_down_generic(const ConvArgs a) {
    const int co = blockIdx.y;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const int KK = a.kh * a.kw;
    EpiCh ech = epi_load_channel(a.epi, co, true);
    float s1[1] = {0.f}, s2[1] = {0.f};
    for (int g0 = blockIdx.x * CAE_NT; g0 < a.total; g0 += gridDim.x * CAE_NT) {
        const int g = g0 + threadIdx.x;
        if (g < a.total) {
            int n = g / (a.QH * a.QW);
            int r = g - n * (a.QH * a.QW);
            int oy = r / a.QW, ox = r - oy * a.QW;
            float acc = 0.f;
            for (int ci = 0; ci < a.Cin; ++ci) {
                const ChanCoef kc = load_coef(a.in, ci);
                const float* wc = a.w + ((size_t)co * a.Cin + ci) * KK;
                for (int ky = 0; ky < a.kh; ++ky) {
                    int iy = oy * a.s + ky - a.p;
                    if (iy < 0 || iy >= iv.H) continue;
                    for (int kx = 0; kx < a.kw; ++kx) {
                        int ix = ox * a.s + kx - a.p;
                        if (ix < 0 || ix >= iv.W) continue;
                        long long off = in_base + (long long)n * iv.sN + (long long)ci * iv.sC + (long long)iy * iv.ld + ix;
                        acc = fmaf(src_value(a.in, off, kc), __ldg(wc + ky * a.kw + kx), acc);
                    }
                }
            }
            epi_element(a.epi, a.out, ech, n, co, oy, ox, acc, 0ll, a.inv_count, s1[0], s2[0]);
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<1>(a.epi, a.out, co, s1, s2);
}

// elementwise member: out = epilogue(in); grid.y = channel
static __global__ void __launch_bounds__(CAE_NT) k_ew_epilogue(const ConvArgs a) {
    const int co = blockIdx.y;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    EpiCh ech = epi_load_channel(a.epi, co, true);
    const ChanCoef kc = load_coef(a.in, co);
    float s1[1] = {0.f}, s2[1] = {0.f};
    for (int g0 = blockIdx.x * CAE_NT; g0 < a.total; g0 += gridDim.x * CAE_NT) {
        const int g = g0 + threadIdx.x;
        if (g < a.total) {
            int n = g / (a.QH * a.QW);
            int r = g - n * (a.QH * a.QW);
            int oy = r / a.QW, ox = r - oy * a.QW;
            long long off = in_base + (long long)n * iv.sN + (long long)co * iv.sC + (long long)oy * iv.ld + ox;
            ChanCoef kk = kc;
            if (a.in.kn) kk.k0 *= __ldg(a.in.kn + (size_t)n * iv.C + co);
            float v = src_value(a.in, off, kk);
            epi_element(a.epi, a.out, ech, n, co, oy, ox, v, tgt_base, a.inv_count, s1[0], s2[0]);
        }
    }
    if (epi_reduces(a.epi.mode)) epi_reduce_tail<1>(a.epi, a.out, co, s1, s2);
}

// elementwise member for [rows, channels] tensors (H = W = 1, channels contiguous: BatchNorm1d of the fc layers).  The plane
// kernel above gives every channel its own CTAs and reads one 4-byte element per 32-byte sector - 76 us for 256 x 3200
// (measured, unet fc 3200); here a CTA owns 32 adjacent channels (128-byte rows), 8 row lanes per channel.
static __global__ void __launch_bounds__(CAE_NT) k_ew_rows(const ConvArgs a) {
    __shared__ float s_part[8][32][2];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int co = blockIdx.x * 32 + cl;
    const bool live = co < a.Cout;
    const CaeView& iv = a.in.t0;
    const long long in_base = src_cursor_offset(a.in);
    const long long tgt_base = (a.epi.mode == CAE_EPI_SIGMOID_MSE) ? src_cursor_offset(a.epi.target) : 0ll;
    const int cc = live ? co : a.Cout - 1;
    EpiCh ech = epi_load_channel(a.epi, cc, live);
    const ChanCoef kc = load_coef(a.in, cc);
    float s1 = 0.f, s2 = 0.f;
    if (live) {
        // four rows per pass: the loads are issued before the first store (in-place calls alias `in` and `out`, which keeps
        // the compiler from hoisting them itself)
        for (int n0 = rl; n0 < a.total; n0 += 32) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = n0 + 8 * q;
                v[q] = 0.f;
                if (n < a.total) {
                    ChanCoef kk = kc;
                    if (a.in.kn) kk.k0 *= __ldg(a.in.kn + (size_t)n * iv.C + co);
                    v[q] = src_value(a.in, in_base + (long long)n * iv.sN + co, kk);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = n0 + 8 * q;
                if (n < a.total) epi_element(a.epi, a.out, ech, n, co, 0, 0, v[q], tgt_base, a.inv_count, s1, s2);
            }
        }
    }
    if (epi_reduces(a.epi.mode)) {
        // the CTA holds every row of its 32 channels: it finishes them itself (no partial rows, no last-CTA pass - the
        // single finalising CTA of the plane kernels walks 3200 channels in 13 dependent passes)
        s_part[rl][cl][0] = s1; s_part[rl][cl][1] = s2;
        __syncthreads();
        if (threadIdx.x < 32 && live) {
            double S = 0.0, Q = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) { S += (double)s_part[r][cl][0]; Q += (double)s_part[r][cl][1]; }
            const double count = (double)a.out.N;
            if (a.epi.mode == CAE_EPI_STATS) bn_forward_channel(a.epi.bn, co, S, Q, count);
            else bn_backward_channel(a.epi.bn, co, S, Q, count);
        }
        if (a.epi.mode == CAE_EPI_STATS && blockIdx.x == 0 && threadIdx.x == 0 && a.epi.bn.num_batches_tracked)
            a.epi.bn.num_batches_tracked[0] += 1;
    }
}

// =======================================================================================
// WGRAD: G[cs][cb][ky][kx] = sum_{n,i,j} small(n,cs,i,j) * big(n,cb,i*S+ky-p,j*S+kx-p)
// One warp owns a CST x CBT tile of (cs,cb) pairs (all taps in registers); lanes stride over the
// positions of this CTA's chunk; warp-shuffle reduction; per-CTA partial rows; last CTA sums rows.
// =======================================================================================
struct WgradArgs {
    CaeSrc sm, bg;
    int kh, kw, s, p;
    float* grad;
    float* partials;
    unsigned int* ticket;
    int Cs, Cb;
    int total;      // N * Hs * Ws
    int chunk;      // positions per chunk
    int nchunks;    // position chunks = rows of `partials`
    int tiles_b;    // ceil(Cb / CBT)
    int ntiles;     // tiles_s * tiles_b ; warp W handles tile W % ntiles of chunk W / ntiles
};

__device__ __forceinline__ void wgrad_final_sum(const WgradArgs& a, int nelem) {
    if (cae_last_block(a.ticket)) {
        const int rows = a.nchunks;
        for (int e = threadIdx.x; e < nelem; e += CAE_NT) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += __ldcg(a.partials + (size_t)r * nelem + e);
            a.grad[e] = s;
        }
    }
}

template <int KH, int KW, int S, int CST, int CBT>
__global__ void __launch_bounds__(CAE_NT) k_conv_wgrad(const WgradArgs a) {
    constexpr int KK = KH * KW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * CAE_NWARP + warp;
    const bool active = wid < a.ntiles * a.nchunks;
    const int wt = wid % a.ntiles, pc = wid / a.ntiles;
    const int cs0 = (wt / a.tiles_b) * CST;
    const int cb0 = (wt % a.tiles_b) * CBT;
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    const int Hs = sv.H, Ws = sv.W, Hb = bv.H, Wb = bv.W;

    float acc[CST][CBT][KK];
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) acc[x][y][t] = 0.f;

    if (active) {
        ChanCoef ks[CST], kb[CBT];
#pragma unroll
        for (int x = 0; x < CST; ++x) ks[x] = load_coef(a.sm, min(cs0 + x, a.Cs - 1));
#pragma unroll
        for (int y = 0; y < CBT; ++y) kb[y] = load_coef(a.bg, min(cb0 + y, a.Cb - 1));
        const int p_begin = pc * a.chunk;
        const int p_end = min(a.total, p_begin + a.chunk);
        for (int g = p_begin + lane; g < p_end; g += 32) {
            int n = g / (Hs * Ws);
            int r = g - n * (Hs * Ws);
            int i = r / Ws, j = r - i * Ws;
            float svv[CST];
            const long long so = sbase + (long long)n * sv.sN + (long long)i * sv.ld + j;
#pragma unroll
            for (int x = 0; x < CST; ++x)
                svv[x] = (cs0 + x < a.Cs) ? src_value(a.sm, so + (long long)(cs0 + x) * sv.sC, ks[x]) : 0.f;
            const int y0 = i * S - a.p, x0 = j * S - a.p;
            const long long bo = bbase + (long long)n * bv.sN + (long long)y0 * bv.ld + x0;
#pragma unroll
            for (int y = 0; y < CBT; ++y) {
                if (cb0 + y < a.Cb) {
                    const long long bco = bo + (long long)(cb0 + y) * bv.sC;
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky)
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx) {
                            const bool ok = (y0 + ky) >= 0 && (y0 + ky) < Hb && (x0 + kx) >= 0 && (x0 + kx) < Wb;
                            float b = ok ? src_value(a.bg, bco + (long long)ky * bv.ld + kx, kb[y]) : 0.f;
#pragma unroll
                            for (int x = 0; x < CST; ++x) acc[x][y][ky * KW + kx] = fmaf(svv[x], b, acc[x][y][ky * KW + kx]);
                        }
                }
            }
        }
    }
    const int nelem = a.Cs * a.Cb * KK;
#pragma unroll
    for (int x = 0; x < CST; ++x)
#pragma unroll
        for (int y = 0; y < CBT; ++y)
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                float v = warp_sum(acc[x][y][t]);
                if (active && lane == 0 && cs0 + x < a.Cs && cb0 + y < a.Cb)
                    a.partials[(size_t)pc * nelem + ((size_t)(cs0 + x) * a.Cb + (cb0 + y)) * KK + t] = v;
            }
    wgrad_final_sum(a, nelem);
}

// generic geometry: one warp per output element (cs, cb, ky, kx)
static __global__ void __launch_bounds__(CAE_NT) k_conv_wgrad_generic(const WgradArgs a) {
    const int KK = a.kh * a.kw;
    const int nelem = a.Cs * a.Cb * KK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * CAE_NWARP + warp;
    const bool active = wid < nelem * a.nchunks;
    const int e = wid % nelem, pc = wid / nelem;
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    float acc = 0.f;
    if (active) {
        const int t = e % KK, cb = (e / KK) % a.Cb, cs = e / (KK * a.Cb);
        const int ky = t / a.kw, kx = t - ky * a.kw;
        const ChanCoef ks = load_coef(a.sm, cs), kb = load_coef(a.bg, cb);
        const int p_begin = pc * a.chunk;
        const int p_end = min(a.total, p_begin + a.chunk);
        for (int g = p_begin + lane; g < p_end; g += 32) {
            int n = g / (sv.H * sv.W);
            int r = g - n * (sv.H * sv.W);
            int i = r / sv.W, j = r - i * sv.W;
            int y = i * a.s + ky - a.p, x = j * a.s + kx - a.p;
            if (y < 0 || y >= bv.H || x < 0 || x >= bv.W) continue;
            float s = src_value(a.sm, sbase + (long long)n * sv.sN + (long long)cs * sv.sC + (long long)i * sv.ld + j, ks);
            float b = src_value(a.bg, bbase + (long long)n * bv.sN + (long long)cb * bv.sC + (long long)y * bv.ld + x, kb);
            acc = fmaf(s, b, acc);
        }
    }
    acc = warp_sum(acc);
    if (active && lane == 0) a.partials[(size_t)pc * nelem + e] = acc;
    wgrad_final_sum(a, nelem);
}

// small problems (few positions, many weight elements - the deep 2x2 ... 8x8 layers): one warp per output element, lanes
// over ALL positions, fixed-order butterfly, direct store.  No partial rows, no ticket, no last-CTA pass: the templated
// kernels above spend 20-35 us on their staging / reduction machinery for < 10 MFLOP of work.
static __global__ void __launch_bounds__(CAE_NT) k_wgrad_small(const WgradArgs a) {
    const int KK = a.kh * a.kw;
    const int nelem = a.Cs * a.Cb * KK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * CAE_NWARP + warp;
    if (e >= nelem) return;
    const CaeView& sv = a.sm.t0;
    const CaeView& bv = a.bg.t0;
    const long long sbase = src_cursor_offset(a.sm), bbase = src_cursor_offset(a.bg);
    const int t = e % KK, cb = (e / KK) % a.Cb, cs = e / (KK * a.Cb);
    const int ky = t / a.kw, kx = t - ky * a.kw;
    const ChanCoef ks = load_coef(a.sm, cs), kb = load_coef(a.bg, cb);
    const int HW = sv.H * sv.W;
    float acc0 = 0.f, acc1 = 0.f;
    int g = lane;
    for (; g + 32 < a.total; g += 64) {                      // two independent positions in flight per lane
        const int g1 = g + 32;
        const int n0 = g / HW, r0 = g - n0 * HW, i0 = r0 / sv.W, j0 = r0 - i0 * sv.W;
        const int n1 = g1 / HW, r1 = g1 - n1 * HW, i1 = r1 / sv.W, j1 = r1 - i1 * sv.W;
        const int y0 = i0 * a.s + ky - a.p, x0 = j0 * a.s + kx - a.p, y1 = i1 * a.s + ky - a.p, x1 = j1 * a.s + kx - a.p;
        const bool ok0 = y0 >= 0 && y0 < bv.H && x0 >= 0 && x0 < bv.W, ok1 = y1 >= 0 && y1 < bv.H && x1 >= 0 && x1 < bv.W;
        float s0 = 0.f, b0 = 0.f, s1 = 0.f, b1 = 0.f;
        if (ok0) {
            s0 = src_value(a.sm, sbase + (long long)n0 * sv.sN + (long long)cs * sv.sC + (long long)i0 * sv.ld + j0, ks);
            b0 = src_value(a.bg, bbase + (long long)n0 * bv.sN + (long long)cb * bv.sC + (long long)y0 * bv.ld + x0, kb);
        }
        if (ok1) {
            s1 = src_value(a.sm, sbase + (long long)n1 * sv.sN + (long long)cs * sv.sC + (long long)i1 * sv.ld + j1, ks);
            b1 = src_value(a.bg, bbase + (long long)n1 * bv.sN + (long long)cb * bv.sC + (long long)y1 * bv.ld + x1, kb);
        }
        acc0 = fmaf(s0, b0, acc0);
        acc1 = fmaf(s1, b1, acc1);
    }
    for (; g < a.total; g += 32) {
        const int n0 = g / HW, r0 = g - n0 * HW, i0 = r0 / sv.W, j0 = r0 - i0 * sv.W;
        const int y0 = i0 * a.s + ky - a.p, x0 = j0 * a.s + kx - a.p;
        if (y0 >= 0 && y0 < bv.H && x0 >= 0 && x0 < bv.W) {
            const float s0 = src_value(a.sm, sbase + (long long)n0 * sv.sN + (long long)cs * sv.sC + (long long)i0 * sv.ld + j0, ks);
            const float b0 = src_value(a.bg, bbase + (long long)n0 * bv.sN + (long long)cb * bv.sC + (long long)y0 * bv.ld + x0, kb);
            acc0 = fmaf(s0, b0, acc0);
        }
    }
    const float acc = warp_sum(acc0 + acc1);
    if (lane == 0) a.grad[e] = acc;
}
