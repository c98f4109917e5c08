// Eval-mode "stem" of the UNET (everything before the last transposed convolution) as ONE kernel:
//   encoder  [Conv2d + BatchNorm2d(eval) + ReLU] x n_conv          (unet.py:73-112)
//   fc stack Linear - BatchNorm1d(eval) - ReLU - Linear - ReLU, twice  (unet.py:92-100,121-129)
//   decoder  [ConvTranspose2d, ChannelAttention gate, concat with the encoder skip, BatchNorm2d(eval), ReLU] x n_up
//            (unet.py:23-39,131-163)
// In eval mode BatchNorm is a per-channel affine, so nothing couples the samples of a batch: a CTA takes STEM_ST = 8 samples
// (two thread halves x 4 samples in registers) through the whole stem with every activation in shared memory (<= 3 K floats
// per sample for the shipped spec); the weights of all layers stay in shared memory for the whole kernel when they fit
// (89 KB for the shipped spec; otherwise they are staged layer by layer) and only the final activated concat tensor is
// written to global memory.  Replaces the 11 launches before the head.
// Measured (B200, stem alone at batch 4096): 577 us with run-time tap loops (round 1) -> ~230 us with compile-time-K loops
// (stem_conv_k / stem_up_s2: ~14 instructions per (channel, tap) for 4 FMAs, issue-bound at 16 warps per SM); the per-layer
// chain needs ~520 us, so the engine uses this kernel at every batch size.
#include "capi_host.h"

#define STEM_S 4             // samples per thread (register block)
#define STEM_H 2             // thread halves: half h takes samples h*STEM_S .. of the pass through the conv / fc / up loops
#define STEM_ST (STEM_S * STEM_H)   // samples per CTA pass
#define STEM_NT 512
#define STEM_HT (STEM_NT / STEM_H)  // threads per half
#define STEM_SMEM_MAX ((size_t)200 * 1024)

struct StemSmem {            // offsets in floats, per CTA (all STEM_ST samples)
    int in0, enc[CAE_STEM_MAX], va, vb, y, cat, small, wbuf;
    int resident;            // 1: every layer's weights live in shared memory for the whole kernel (offsets below), 0: staged per layer
    int w_conv[CAE_STEM_MAX], w_fc[CAE_STEM_MAX], w_up[CAE_STEM_MAX];
    int total;
};

struct StemArgs {
    CaeUnetStem s;
    CaeSrc x;
    CaeView out;
    StemSmem m;
};

__device__ __forceinline__ float stem_bn_relu(float v, const float* scale, const float* shift, int c) {
    if (scale) v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
    return fmaxf(v, 0.f);
}

// y[s][e] for e = (co, oy, ox): strided convolution with padding, BN(eval) + ReLU
__device__ __forceinline__ void stem_conv(const CaeStemConv& L, const float* wsm, const float* in, float* outp, int in_pitch,
                                          int out_pitch) {
    const int HWo = L.Hout * L.Wout, total = L.Cout * HWo, KK = L.k * L.k;
    for (int e = (threadIdx.x & (STEM_HT - 1)); e < total; e += STEM_HT) {
        const int co = e / HWo, r = e - co * HWo, oy = r / L.Wout, ox = r - oy * L.Wout;
        float acc[STEM_S];
        const float b = L.b ? __ldg(L.b + co) : 0.f;
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) acc[s] = b;
        const int iy0 = oy * L.stride - L.pad, ix0 = ox * L.stride - L.pad;
        for (int ci = 0; ci < L.Cin; ++ci) {
            const float* wp = wsm + (co * L.Cin + ci) * KK;
            const float* ip = in + ci * L.Hin * L.Win;
            for (int ky = 0; ky < L.k; ++ky) {
                const int iy = iy0 + ky;
                if (iy < 0 || iy >= L.Hin) continue;
                for (int kx = 0; kx < L.k; ++kx) {
                    const int ix = ix0 + kx;
                    if (ix < 0 || ix >= L.Win) continue;
                    const float wv = wp[ky * L.k + kx];
                    const float* q = ip + iy * L.Win + ix;
#pragma unroll
                    for (int s = 0; s < STEM_S; ++s) acc[s] = fmaf(q[s * in_pitch], wv, acc[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) outp[s * out_pitch + e] = stem_bn_relu(acc[s], L.scale, L.shift, co);
    }
}

// Compile-time kernel size: the K*K taps are unrolled, their bounds checks become per-thread masks computed once per layer, and
// the 9 / 16 independent (weight, S inputs) load groups of one input channel are in flight together.  The generic loop above
// re-evaluates the bounds inside a run-time triple loop: one dependent shared-memory round trip per FMA group (~70 cycles per
// tap measured through the whole kernel: 577 us per 4096 samples).
template <int K>
__device__ __forceinline__ void stem_conv_k(const CaeStemConv& L, const float* wsm, const float* in, float* outp, int in_pitch,
                                            int out_pitch) {
    constexpr int KK = K * K;
    const int HWo = L.Hout * L.Wout, total = L.Cout * HWo, HWi = L.Hin * L.Win;
    for (int e = (threadIdx.x & (STEM_HT - 1)); e < total; e += STEM_HT) {
        const int co = e / HWo, r = e - co * HWo, oy = r / L.Wout, ox = r - oy * L.Wout;
        const int iy0 = oy * L.stride - L.pad, ix0 = ox * L.stride - L.pad;
        int off[KK];
        bool ok[KK];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int iy = iy0 + ky, ix = ix0 + kx;
                ok[ky * K + kx] = iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win;
                off[ky * K + kx] = ok[ky * K + kx] ? iy * L.Win + ix : 0;
            }
        float acc[STEM_S];
        const float b = L.b ? __ldg(L.b + co) : 0.f;
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) acc[s] = b;
        const float* wp = wsm + co * L.Cin * KK;
#pragma unroll 2
        for (int ci = 0; ci < L.Cin; ++ci) {
            const float* ip = in + ci * HWi;
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                const float wv = ok[t] ? wp[ci * KK + t] : 0.f;
                const float* q = ip + off[t];
#pragma unroll
                for (int s = 0; s < STEM_S; ++s) acc[s] = fmaf(q[s * in_pitch], wv, acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) outp[s * out_pitch + e] = stem_bn_relu(acc[s], L.scale, L.shift, co);
    }
}

// transposed convolution with stride 2 and k <= 4: every output pixel has at most 2 x 2 taps (ky = (oy + pad) mod 2 + 2a);
// they are located once per layer and thread - the generic gather below pays a modulo and a division per tap and channel
__device__ __forceinline__ void stem_up_s2(const CaeStemUp& L, const float* wsm, const float* in, float* outp, int in_pitch,
                                           int out_pitch) {
    const int HWo = L.Hout * L.Wout, total = L.Cout * HWo, KK = L.k * L.k, HWi = L.Hin * L.Win, wci = L.Cout * KK;
    for (int e = (threadIdx.x & (STEM_HT - 1)); e < total; e += STEM_HT) {
        const int co = e / HWo, r = e - co * HWo, oy = r / L.Wout, ox = r - oy * L.Wout;
        int pos[4], tap[4];
        bool ok[4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                const int ky = ((oy + L.pad) & 1) + 2 * a, kx = ((ox + L.pad) & 1) + 2 * b2;
                const int iy = ((oy + L.pad) >> 1) - a, ix = ((ox + L.pad) >> 1) - b2;
                const bool v = ky < L.k && kx < L.k && iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win;
                ok[a * 2 + b2] = v;
                pos[a * 2 + b2] = v ? iy * L.Win + ix : 0;
                tap[a * 2 + b2] = v ? co * KK + ky * L.k + kx : 0;
            }
        float acc[STEM_S];
        const float b = L.b ? __ldg(L.b + co) : 0.f;
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) acc[s] = b;
#pragma unroll 2
        for (int ci = 0; ci < L.Cin; ++ci) {
            const float* wp = wsm + ci * wci;
            const float* ip = in + ci * HWi;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float wv = ok[t] ? wp[tap[t]] : 0.f;
                const float* q = ip + pos[t];
#pragma unroll
                for (int s = 0; s < STEM_S; ++s) acc[s] = fmaf(q[s * in_pitch], wv, acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) outp[s * out_pitch + e] = acc[s];
    }
}

// v_out[s][o] = act((W v_in[s] + b) * scale + shift)
__device__ __forceinline__ void stem_fc(const CaeStemFc& L, const float* wsm, const float* in, float* outp, int in_pitch,
                                        int out_pitch) {
    const int ht = threadIdx.x & (STEM_HT - 1), lane = ht & 31, warp = ht >> 5;
    if (L.in >= 64) {                                   // long rows: one warp per output, lanes over k
        for (int o = warp; o < L.out; o += STEM_HT / 32) {
            float acc[STEM_S];
#pragma unroll
            for (int s = 0; s < STEM_S; ++s) acc[s] = 0.f;
            for (int k = lane; k < L.in; k += 32) {
                const float wv = wsm[o * L.in + k];
#pragma unroll
                for (int s = 0; s < STEM_S; ++s) acc[s] = fmaf(in[s * in_pitch + k], wv, acc[s]);
            }
#pragma unroll
            for (int s = 0; s < STEM_S; ++s) acc[s] = warp_sum(acc[s]);
            if (lane < STEM_S) {
                float v = acc[0];
#pragma unroll
                for (int s = 1; s < STEM_S; ++s) if (lane == s) v = acc[s];
                v += L.b ? __ldg(L.b + o) : 0.f;
                if (L.scale) v = fmaf(v, __ldg(L.scale + o), __ldg(L.shift + o));
                outp[lane * out_pitch + o] = L.relu ? fmaxf(v, 0.f) : v;
            }
        }
    } else {
        for (int o = ht; o < L.out; o += STEM_HT) {
            float acc[STEM_S];
            const float b = L.b ? __ldg(L.b + o) : 0.f;
#pragma unroll
            for (int s = 0; s < STEM_S; ++s) acc[s] = b;
            for (int k = 0; k < L.in; ++k) {
                const float wv = wsm[o * L.in + k];
#pragma unroll
                for (int s = 0; s < STEM_S; ++s) acc[s] = fmaf(in[s * in_pitch + k], wv, acc[s]);
            }
#pragma unroll
            for (int s = 0; s < STEM_S; ++s) {
                float v = acc[s];
                if (L.scale) v = fmaf(v, __ldg(L.scale + o), __ldg(L.shift + o));
                outp[s * out_pitch + o] = L.relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
}

// transposed convolution (gather form, padding), raw output + bias
__device__ __forceinline__ void stem_up(const CaeStemUp& L, const float* wsm, const float* in, float* outp, int in_pitch,
                                        int out_pitch) {
    const int HWo = L.Hout * L.Wout, total = L.Cout * HWo, KK = L.k * L.k;
    for (int e = (threadIdx.x & (STEM_HT - 1)); e < total; e += STEM_HT) {
        const int co = e / HWo, r = e - co * HWo, oy = r / L.Wout, ox = r - oy * L.Wout;
        float acc[STEM_S];
        const float b = L.b ? __ldg(L.b + co) : 0.f;
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) acc[s] = b;
        for (int ky = 0; ky < L.k; ++ky) {
            const int ty = oy + L.pad - ky;
            if (ty < 0 || ty % L.stride) continue;
            const int iy = ty / L.stride;
            if (iy >= L.Hin) continue;
            for (int kx = 0; kx < L.k; ++kx) {
                const int tx = ox + L.pad - kx;
                if (tx < 0 || tx % L.stride) continue;
                const int ix = tx / L.stride;
                if (ix >= L.Win) continue;
                const float* wp = wsm + co * KK + ky * L.k + kx;                    // + ci * Cout * KK
                const float* q = in + iy * L.Win + ix;                              // + ci * Hin * Win
#pragma unroll 4
                for (int ci = 0; ci < L.Cin; ++ci) {
                    const float wv = wp[ci * L.Cout * KK];
#pragma unroll
                    for (int s = 0; s < STEM_S; ++s) acc[s] = fmaf(q[s * in_pitch + ci * L.Hin * L.Win], wv, acc[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < STEM_S; ++s) outp[s * out_pitch + e] = acc[s];
    }
}

__global__ void __launch_bounds__(STEM_NT) k_unet_stem_eval(const StemArgs a) {
    extern __shared__ __align__(16) float sm[];
    const CaeUnetStem& S = a.s;
    const StemSmem& m = a.m;
    const int tid = threadIdx.x;
    const CaeStemConv& c0 = S.conv[0];
    const int in_elems = c0.Cin * c0.Hin * c0.Win;
    const CaeView& xv = a.x.t0;
    const long long xbase = src_cursor_offset(a.x);
    const int N = xv.N;
    float* wsm = sm + m.wbuf;
    const int h = tid / STEM_HT;                           // thread half -> samples h * STEM_S ..
    // Layer weights go through shared memory (with the SM carved out for activations the L1 keeps almost nothing, and weight
    // loads that miss it cost an L2 round trip per tap).  When everything fits they are staged ONCE per CTA (`resident`);
    // otherwise layer by layer into one buffer, as the first version of this kernel did for every pass of 4 samples.
    auto stage = [&](float* dst, const float* w, int n) {
        for (int e = tid; e < n; e += STEM_NT) dst[e] = __ldg(w + e);
    };
    if (m.resident) {
        for (int l = 0; l < S.n_conv; ++l) stage(sm + m.w_conv[l], S.conv[l].w, S.conv[l].Cout * S.conv[l].Cin * S.conv[l].k * S.conv[l].k);
        for (int l = 0; l < S.n_fc; ++l) stage(sm + m.w_fc[l], S.fc[l].w, S.fc[l].in * S.fc[l].out);
        for (int j = 0; j < S.n_up; ++j) stage(sm + m.w_up[j], S.up[j].w, S.up[j].Cin * S.up[j].Cout * S.up[j].k * S.up[j].k);
    }
    auto weights = [&](int off, const float* w, int n) -> const float* {
        if (m.resident) return sm + off;
        stage(wsm, w, n);
        __syncthreads();
        return wsm;
    };
    for (int n0 = blockIdx.x * STEM_ST; n0 < N; n0 += gridDim.x * STEM_ST) {
        __syncthreads();
        // ---- input (samples beyond N are zero-filled; their results are never written)
        for (int e = tid; e < STEM_ST * in_elems; e += STEM_NT) {
            const int s = e / in_elems, r = e - s * in_elems;
            const int c = r / (c0.Hin * c0.Win), q = r - c * c0.Hin * c0.Win, yy = q / c0.Win, xx = q - yy * c0.Win;
            float v = 0.f;
            if (n0 + s < N) {
                const ChanCoef kc = load_coef(a.x, c);
                v = src_value(a.x, xbase + (long long)(n0 + s) * xv.sN + (long long)c * xv.sC + (long long)yy * xv.ld + xx, kc);
            }
            sm[m.in0 + e] = v;
        }
        __syncthreads();
        // ---- encoder
        const float* cur = sm + m.in0;
        int cur_pitch = in_elems;
        for (int l = 0; l < S.n_conv; ++l) {
            const CaeStemConv& L = S.conv[l];
            const int op = L.Cout * L.Hout * L.Wout;
            const float* w = weights(m.w_conv[l], L.w, L.Cout * L.Cin * L.k * L.k);
            const float* in = cur + h * STEM_S * cur_pitch;
            float* outp = sm + m.enc[l] + h * STEM_S * op;
            if (L.k == 3) stem_conv_k<3>(L, w, in, outp, cur_pitch, op);
            else if (L.k == 4) stem_conv_k<4>(L, w, in, outp, cur_pitch, op);
            else stem_conv(L, w, in, outp, cur_pitch, op);
            __syncthreads();
            cur = sm + m.enc[l];
            cur_pitch = op;
        }
        // ---- fc stack (ping-pong va / vb)
        float* va = sm + m.va;
        float* vb = sm + m.vb;
        for (int l = 0; l < S.n_fc; ++l) {
            const CaeStemFc& L = S.fc[l];
            float* dst = (l & 1) ? vb : va;
            const float* w = weights(m.w_fc[l], L.w, L.in * L.out);
            stem_fc(L, w, cur + h * STEM_S * cur_pitch, dst + h * STEM_S * L.out, cur_pitch, L.out);
            __syncthreads();
            cur = dst;
            cur_pitch = L.out;
        }
        // ---- decoder blocks
        for (int j = 0; j < S.n_up; ++j) {
            const CaeStemUp& L = S.up[j];
            const int C = L.Cout, HW = L.Hout * L.Wout, yp = C * HW, cp = 2 * yp;
            float* y = sm + m.y;
            float* cat = sm + m.cat;
            const float* w = weights(m.w_up[j], L.w, L.Cin * L.Cout * L.k * L.k);
            if (L.stride == 2 && L.k <= 4) stem_up_s2(L, w, cur + h * STEM_S * cur_pitch, y + h * STEM_S * yp, cur_pitch, yp);
            else stem_up(L, w, cur + h * STEM_S * cur_pitch, y + h * STEM_S * yp, cur_pitch, yp);
            __syncthreads();
            float* avg = sm + m.small;                     // [ST][C]
            float* mx = avg + STEM_ST * C;                 // [ST][C]
            float* hid = mx + STEM_ST * C;                 // [ST][2][Cr]
            float* att = hid + STEM_ST * 2 * L.Cr;         // [ST][C]
            for (int e = tid; e < STEM_ST * C; e += STEM_NT) {
                const int s = e / C, c = e - s * C;
                const float* q = y + s * yp + c * HW;
                float sum = 0.f, mxx = -INFINITY;
                for (int i = 0; i < HW; ++i) { sum += q[i]; mxx = fmaxf(mxx, q[i]); }
                avg[e] = sum / (float)HW;
                mx[e] = mxx;
            }
            __syncthreads();
            for (int e = tid; e < STEM_ST * 2 * L.Cr; e += STEM_NT) {
                const int s = e / (2 * L.Cr), r2 = e - s * 2 * L.Cr, which = r2 / L.Cr, r = r2 - which * L.Cr;
                const float* src = (which ? mx : avg) + s * C;
                float v = 0.f;
                for (int c = 0; c < C; ++c) v = fmaf(__ldg(L.W1 + r * C + c), src[c], v);
                hid[e] = fmaxf(v, 0.f);
            }
            __syncthreads();
            for (int e = tid; e < STEM_ST * C; e += STEM_NT) {
                const int s = e / C, c = e - s * C;
                const float* hd = hid + s * 2 * L.Cr;
                float v = 0.f;
                for (int r = 0; r < L.Cr; ++r) v = fmaf(__ldg(L.W2 + c * L.Cr + r), hd[r] + hd[L.Cr + r], v);
                att[e] = 1.f / (1.f + expf(-v));
            }
            __syncthreads();
            // cat = relu(bn([att * y ; skip]))
            const float* skip = sm + m.enc[L.skip];
            const bool last = j == S.n_up - 1;
            for (int e = tid; e < STEM_ST * cp; e += STEM_NT) {
                const int s = e / cp, r = e - s * cp, c2 = r / HW, i = r - c2 * HW;
                float v = c2 < C ? att[s * C + c2] * y[s * yp + r] : skip[s * yp + (r - yp)];
                v = stem_bn_relu(v, L.scale, L.shift, c2);
                cat[e] = v;
                if (last && n0 + s < N) {
                    const int yy = i / L.Wout, xx = i - yy * L.Wout;
                    a.out.p[(long long)(n0 + s) * a.out.sN + (long long)c2 * a.out.sC + (long long)yy * a.out.ld + xx] = v;
                }
            }
            __syncthreads();
            cur = cat;
            cur_pitch = cp;
        }
    }
}

static int stem_plan(const CaeUnetStem& s, StemSmem& m, char* why, size_t why_len) {
#define STEM_FAIL(...) do { snprintf(why, why_len, __VA_ARGS__); return 0; } while (0)
    if (s.n_conv < 1 || s.n_conv > CAE_STEM_MAX || s.n_fc < 1 || s.n_fc > CAE_STEM_MAX || s.n_up < 1 || s.n_up > CAE_STEM_MAX)
        STEM_FAIL("layer counts %d/%d/%d outside 1..%d", s.n_conv, s.n_fc, s.n_up, CAE_STEM_MAX);
    int off = 0;
    auto take = [&](int floats) { int o = off; off += (floats + 3) & ~3; return o; };
    m.in0 = take(STEM_ST * s.conv[0].Cin * s.conv[0].Hin * s.conv[0].Win);
    int prev = s.conv[0].Cin * s.conv[0].Hin * s.conv[0].Win;
    for (int l = 0; l < s.n_conv; ++l) {
        const CaeStemConv& L = s.conv[l];
        if (!L.w || L.Cin * L.Hin * L.Win != prev) STEM_FAIL("encoder layer %d does not chain", l);
        if ((L.Hin + 2 * L.pad - L.k) / L.stride + 1 != L.Hout || (L.Win + 2 * L.pad - L.k) / L.stride + 1 != L.Wout)
            STEM_FAIL("encoder layer %d geometry", l);
        prev = L.Cout * L.Hout * L.Wout;
        m.enc[l] = take(STEM_ST * prev);
    }
    int vmax = 0;
    for (int l = 0; l < s.n_fc; ++l) {
        if (!s.fc[l].w || s.fc[l].in != prev) STEM_FAIL("fc layer %d does not chain (%d vs %d)", l, s.fc[l].in, prev);
        prev = s.fc[l].out;
        vmax = max(vmax, prev);
    }
    m.va = take(STEM_ST * vmax);
    m.vb = take(STEM_ST * vmax);
    int ymax = 0, smallmax = 0;
    for (int j = 0; j < s.n_up; ++j) {
        const CaeStemUp& L = s.up[j];
        if (!L.w || !L.W1 || !L.W2 || L.Cin * L.Hin * L.Win != prev) STEM_FAIL("decoder block %d does not chain", j);
        if ((L.Hin - 1) * L.stride - 2 * L.pad + L.k != L.Hout || (L.Win - 1) * L.stride - 2 * L.pad + L.k != L.Wout)
            STEM_FAIL("decoder block %d geometry (output padding is not supported)", j);
        if (L.skip < 0 || L.skip >= s.n_conv) STEM_FAIL("decoder block %d: bad skip index", j);
        const CaeStemConv& E = s.conv[L.skip];
        if (E.Cout != L.Cout || E.Hout != L.Hout || E.Wout != L.Wout) STEM_FAIL("decoder block %d: skip geometry mismatch", j);
        ymax = max(ymax, L.Cout * L.Hout * L.Wout);
        smallmax = max(smallmax, 3 * L.Cout + 2 * L.Cr);
        prev = 2 * L.Cout * L.Hout * L.Wout;
    }
    m.y = take(STEM_ST * ymax);
    m.cat = take(STEM_ST * 2 * ymax);
    m.small = take(STEM_ST * smallmax);
    int wmax = 0, wsum = 0;
    auto wsize = [&](int n) { wmax = max(wmax, n); wsum += (n + 3) & ~3; };
    for (int l = 0; l < s.n_conv; ++l) wsize(s.conv[l].Cout * s.conv[l].Cin * s.conv[l].k * s.conv[l].k);
    for (int l = 0; l < s.n_fc; ++l) wsize(s.fc[l].in * s.fc[l].out);
    for (int j = 0; j < s.n_up; ++j) wsize(s.up[j].Cin * s.up[j].Cout * s.up[j].k * s.up[j].k);
    m.resident = (size_t)(off + wsum) * 4 <= STEM_SMEM_MAX;
    if (m.resident) {
        for (int l = 0; l < s.n_conv; ++l) m.w_conv[l] = take(s.conv[l].Cout * s.conv[l].Cin * s.conv[l].k * s.conv[l].k);
        for (int l = 0; l < s.n_fc; ++l) m.w_fc[l] = take(s.fc[l].in * s.fc[l].out);
        for (int j = 0; j < s.n_up; ++j) m.w_up[j] = take(s.up[j].Cin * s.up[j].Cout * s.up[j].k * s.up[j].k);
        m.wbuf = off;
    } else {
        m.wbuf = take(wmax);
    }
    m.total = off;
    if ((size_t)off * 4 > STEM_SMEM_MAX)
        STEM_FAIL("activations of %d samples need %d KB of shared memory (> %d)", STEM_ST, off * 4 / 1024, (int)(STEM_SMEM_MAX / 1024));
    return 1;
#undef STEM_FAIL
}

extern "C" int cae_unet_stem_supported(const CaeUnetStem* s) {
    if (!s) return 0;
    StemSmem m;
    char why[128];
    return stem_plan(*s, m, why, sizeof(why));
}

extern "C" int cae_unet_stem_eval(const CaeUnetStem* s, const CaeSrc* x, const CaeView* out, void* stream) {
    CAE_REQUIRE(s && x && out, "unet_stem_eval: null argument");
    int rc;
    if ((rc = check_view(x->t0, "unet_stem_eval input"))) return rc;
    if ((rc = check_view(*out, "unet_stem_eval output"))) return rc;
    StemArgs a;
    memset(&a, 0, sizeof(a));
    char why[128] = "";
    if (!stem_plan(*s, a.m, why, sizeof(why))) {
        cae_set_error("unet_stem_eval: %s", why);
        return CAE_EUNSUPPORTED;
    }
    const CaeStemConv& c0 = s->conv[0];
    const CaeStemUp& ul = s->up[s->n_up - 1];
    CAE_REQUIRE(x->t0.C == c0.Cin && x->t0.H == c0.Hin && x->t0.W == c0.Win, "unet_stem_eval: input geometry differs from the first layer");
    CAE_REQUIRE(out->N == x->t0.N && out->C == 2 * ul.Cout && out->H == ul.Hout && out->W == ul.Wout,
                "unet_stem_eval: output must be [N, %d, %d, %d]", 2 * ul.Cout, ul.Hout, ul.Wout);
    CAE_REQUIRE(x->kn == nullptr, "unet_stem_eval: kn not supported");
    a.s = *s; a.x = *x; a.out = *out;
    const size_t smem = (size_t)a.m.total * 4;
    static bool opted = false;
    if (!opted) {
        cudaFuncSetAttribute(k_unet_stem_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEM_SMEM_MAX);
        opted = true;
    }
    const int passes = ceil_div(x->t0.N, STEM_ST);
    const int per_sm = max(1, min(4, (int)((220 * 1024) / (smem + 1024))));
    const int grid = min(passes, CAE_NUM_SMS * per_sm);
    k_unet_stem_eval<<<grid, STEM_NT, smem, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_unet_stem_eval");
}
