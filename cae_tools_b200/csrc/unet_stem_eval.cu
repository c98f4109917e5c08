// Eval-mode "stem" of the UNET (everything before the last transposed convolution) as ONE kernel:
//   encoder  [Conv2d + BatchNorm2d(eval) + ReLU] x n_conv          (unet.py:73-112)
//   fc stack Linear - BatchNorm1d(eval) - ReLU - Linear - ReLU, twice  (unet.py:92-100,121-129)
//   decoder  [ConvTranspose2d, ChannelAttention gate, concat with the encoder skip, BatchNorm2d(eval), ReLU] x n_up
//            (unet.py:23-39,131-163)
// In eval mode BatchNorm is a per-channel affine, so nothing couples the samples of a batch: a CTA of 512 threads takes
// STEM_ST = 8 samples through the whole stem with every activation in shared memory (<= 3 K floats per sample for the
// shipped spec); the weights of all layers stay in shared memory for the whole kernel when they fit (89 KB for the shipped
// spec; otherwise they are staged layer by layer) and only the final activated concat tensor is written to global memory.
// Replaces the 11 launches before the head.
// Measured (B200, stem alone at batch 4096; tools/eval_stem_probe.py): 577 us with run-time tap loops (round 1) -> 230 us
// with compile-time-K loops, one output x 4 samples per thread (4 bytes of shared-memory traffic per FMA: the loops ran at
// the shared-memory bandwidth, and neither resident weights nor 16-byte loads changed the time) -> 190 us with the current
// item = (sample, position, 4 output channels) mapping, which is issue-bound (~2.3 instructions per FMA plus ~200 of index
// arithmetic per item).  The per-layer chain needs ~520 us, so the engine uses this kernel at every batch size.
#include "capi_host.h"

#define STEM_ST 8            // samples per CTA pass
#define STEM_NT 512
#define STEM_COB 4           // output channels per thread (register block) in the fast conv / up paths
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

// Layouts in shared memory.
//   activations: the STEM_ST samples of the pass are INNERMOST - element e of sample s at base + e * STEM_ST + s;
//   conv / up weights: [ci][tap][co] (re-laid-out while staging), fc weights [out][in] as stored.
// Work item of the conv / up loops = (sample s, output position, group of COB output channels), s fastest: the 8 lanes of
// a position read 8 consecutive floats, the COB weights of a tap arrive as one broadcast 16-byte load, and every input value
// that leaves shared memory feeds COB FMAs.
__device__ __forceinline__ int stem_idx(int s, int e) { return e * STEM_ST + s; }

template <int COB>
__device__ __forceinline__ void stem_ldw(const float* p, float (&w)[COB]) {
    if constexpr (COB == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    } else {
#pragma unroll
        for (int j = 0; j < COB; ++j) w[j] = p[j];
    }
}

// strided convolution with padding, BN(eval) + ReLU; K > 0: compile-time kernel size (taps unrolled, bounds as per-thread masks
// computed once per item), K == 0: any kernel size (run-time loops)
template <int K, int COB>
__device__ __forceinline__ void stem_conv(const CaeStemConv& L, const float* wsm, const float* in, float* outp) {
    const int k = K ? K : L.k, KK = k * k;
    const int HWo = L.Hout * L.Wout, HWi = L.Hin * L.Win, total = (L.Cout / COB) * HWo * STEM_ST;
    for (int i = threadIdx.x; i < total; i += STEM_NT) {
        const int s = i % STEM_ST, rest = i / STEM_ST, pos = rest % HWo, co0 = (rest / HWo) * COB;
        const int oy = pos / L.Wout, ox = pos - oy * L.Wout;
        const int iy0 = oy * L.stride - L.pad, ix0 = ox * L.stride - L.pad;
        float acc[COB];
#pragma unroll
        for (int j = 0; j < COB; ++j) acc[j] = L.b ? __ldg(L.b + co0 + j) : 0.f;
        const float* ip0 = in + s;
        const float* wp0 = wsm + co0;
        if constexpr (K > 0) {
            int off[K * K];
            bool ok[K * K];
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const int iy = iy0 + ky, ix = ix0 + kx;
                    ok[ky * K + kx] = iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win;
                    off[ky * K + kx] = ok[ky * K + kx] ? (iy * L.Win + ix) * STEM_ST : 0;
                }
#pragma unroll 2
            for (int ci = 0; ci < L.Cin; ++ci) {
                const float* ip = ip0 + ci * HWi * STEM_ST;
                const float* wp = wp0 + ci * KK * L.Cout;
#pragma unroll
                for (int t = 0; t < K * K; ++t) {
                    const float x = ok[t] ? ip[off[t]] : 0.f;
                    float w[COB];
                    stem_ldw<COB>(wp + t * L.Cout, w);
#pragma unroll
                    for (int j = 0; j < COB; ++j) acc[j] = fmaf(x, w[j], acc[j]);
                }
            }
        } else {
            for (int ci = 0; ci < L.Cin; ++ci)
                for (int ky = 0; ky < k; ++ky) {
                    const int iy = iy0 + ky;
                    if (iy < 0 || iy >= L.Hin) continue;
                    for (int kx = 0; kx < k; ++kx) {
                        const int ix = ix0 + kx;
                        if (ix < 0 || ix >= L.Win) continue;
                        const float x = ip0[(ci * HWi + iy * L.Win + ix) * STEM_ST];
                        float w[COB];
                        stem_ldw<COB>(wp0 + (ci * KK + ky * k + kx) * L.Cout, w);
#pragma unroll
                        for (int j = 0; j < COB; ++j) acc[j] = fmaf(x, w[j], acc[j]);
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < COB; ++j) outp[stem_idx(s, (co0 + j) * HWo + pos)] = stem_bn_relu(acc[j], L.scale, L.shift, co0 + j);
    }
}

// transposed convolution (gather form, padding), raw output + bias.  S2: stride 2 and k <= 4 - every output pixel has at most
// 2 x 2 taps (ky = (oy + pad) mod 2 + 2a), located once per item; otherwise run-time loops with a modulo per tap.
template <bool S2, int COB>
__device__ __forceinline__ void stem_up(const CaeStemUp& L, const float* wsm, const float* in, float* outp) {
    const int KK = L.k * L.k, HWo = L.Hout * L.Wout, HWi = L.Hin * L.Win, total = (L.Cout / COB) * HWo * STEM_ST;
    for (int i = threadIdx.x; i < total; i += STEM_NT) {
        const int s = i % STEM_ST, rest = i / STEM_ST, pos = rest % HWo, co0 = (rest / HWo) * COB;
        const int oy = pos / L.Wout, ox = pos - oy * L.Wout;
        float acc[COB];
#pragma unroll
        for (int j = 0; j < COB; ++j) acc[j] = L.b ? __ldg(L.b + co0 + j) : 0.f;
        const float* ip0 = in + s;
        const float* wp0 = wsm + co0;
        if constexpr (S2) {
            int off[4], tap[4];
            bool ok[4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    const int ky = ((oy + L.pad) & 1) + 2 * a, kx = ((ox + L.pad) & 1) + 2 * b2;
                    const int iy = ((oy + L.pad) >> 1) - a, ix = ((ox + L.pad) >> 1) - b2;
                    const bool v = ky < L.k && kx < L.k && iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win;
                    ok[a * 2 + b2] = v;
                    off[a * 2 + b2] = v ? (iy * L.Win + ix) * STEM_ST : 0;
                    tap[a * 2 + b2] = v ? (ky * L.k + kx) * L.Cout : 0;
                }
#pragma unroll 2
            for (int ci = 0; ci < L.Cin; ++ci) {
                const float* ip = ip0 + ci * HWi * STEM_ST;
                const float* wp = wp0 + ci * KK * L.Cout;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float x = ok[t] ? ip[off[t]] : 0.f;
                    float w[COB];
                    stem_ldw<COB>(wp + tap[t], w);
#pragma unroll
                    for (int j = 0; j < COB; ++j) acc[j] = fmaf(x, w[j], acc[j]);
                }
            }
        } else {
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
                    for (int ci = 0; ci < L.Cin; ++ci) {
                        const float x = ip0[(ci * HWi + iy * L.Win + ix) * STEM_ST];
                        float w[COB];
                        stem_ldw<COB>(wp0 + (ci * KK + ky * L.k + kx) * L.Cout, w);
#pragma unroll
                        for (int j = 0; j < COB; ++j) acc[j] = fmaf(x, w[j], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < COB; ++j) outp[stem_idx(s, (co0 + j) * HWo + pos)] = acc[j];
    }
}

// v_out[o][s] = act((W v_in[.][s] + b) * scale + shift): item = (s, o)
__device__ __forceinline__ void stem_fc(const CaeStemFc& L, const float* wsm, const float* in, float* outp) {
    for (int i = threadIdx.x; i < L.out * STEM_ST; i += STEM_NT) {
        const int s = i % STEM_ST, o = i / STEM_ST;
        const float* wp = wsm + o * L.in;
        const float* ip = in + s;
        float a0 = L.b ? __ldg(L.b + o) : 0.f, a1 = 0.f;
        int k = 0;
        for (; k + 1 < L.in; k += 2) {
            a0 = fmaf(ip[k * STEM_ST], wp[k], a0);
            a1 = fmaf(ip[(k + 1) * STEM_ST], wp[k + 1], a1);
        }
        if (k < L.in) a0 = fmaf(ip[k * STEM_ST], wp[k], a0);
        float v = a0 + a1;
        if (L.scale) v = fmaf(v, __ldg(L.scale + o), __ldg(L.shift + o));
        outp[stem_idx(s, o)] = L.relu ? fmaxf(v, 0.f) : v;
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
    // Layer weights go through shared memory, re-laid-out to [ci][tap][co] for the conv / up layers.  When everything fits
    // they are staged ONCE per CTA (`resident`); otherwise layer by layer into one buffer.
    // kind 0: conv weights [co][ci][KK]; 1: transposed-conv weights [ci][co][KK]; 2: fc weights (copied as they are)
    auto stage = [&](float* dst, const float* w, int kind, int Cout, int Cin, int KK) {
        const int n = Cout * Cin * KK;
        for (int e = tid; e < n; e += STEM_NT) {
            int d = e;
            if (kind == 0) {
                const int co = e / (Cin * KK), r = e - co * Cin * KK;          // r = ci * KK + t
                d = r * Cout + co;
            } else if (kind == 1) {
                const int ci = e / (Cout * KK), r = e - ci * Cout * KK, co = r / KK, t = r - co * KK;
                d = (ci * KK + t) * Cout + co;
            }
            dst[d] = __ldg(w + e);
        }
    };
    if (m.resident) {
        for (int l = 0; l < S.n_conv; ++l) stage(sm + m.w_conv[l], S.conv[l].w, 0, S.conv[l].Cout, S.conv[l].Cin, S.conv[l].k * S.conv[l].k);
        for (int l = 0; l < S.n_fc; ++l) stage(sm + m.w_fc[l], S.fc[l].w, 2, S.fc[l].out, S.fc[l].in, 1);
        for (int j = 0; j < S.n_up; ++j) stage(sm + m.w_up[j], S.up[j].w, 1, S.up[j].Cout, S.up[j].Cin, S.up[j].k * S.up[j].k);
    }
    auto weights = [&](int off, const float* w, int kind, int Cout, int Cin, int KK) -> const float* {
        if (m.resident) return sm + off;
        stage(wsm, w, kind, Cout, Cin, KK);
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
            sm[m.in0 + stem_idx(s, r)] = v;
        }
        __syncthreads();
        // ---- encoder
        const float* cur = sm + m.in0;
        for (int l = 0; l < S.n_conv; ++l) {
            const CaeStemConv& L = S.conv[l];
            const float* w = weights(m.w_conv[l], L.w, 0, L.Cout, L.Cin, L.k * L.k);
            float* outp = sm + m.enc[l];
            const bool blk = L.Cout % STEM_COB == 0;
            if (L.k == 3 && blk) stem_conv<3, STEM_COB>(L, w, cur, outp);
            else if (L.k == 4 && blk) stem_conv<4, STEM_COB>(L, w, cur, outp);
            else if (L.k == 3) stem_conv<3, 1>(L, w, cur, outp);
            else stem_conv<0, 1>(L, w, cur, outp);
            __syncthreads();
            cur = outp;
        }
        // ---- fc stack (ping-pong va / vb)
        float* va = sm + m.va;
        float* vb = sm + m.vb;
        for (int l = 0; l < S.n_fc; ++l) {
            const CaeStemFc& L = S.fc[l];
            float* dst = (l & 1) ? vb : va;
            const float* w = weights(m.w_fc[l], L.w, 2, L.out, L.in, 1);
            stem_fc(L, w, cur, dst);
            __syncthreads();
            cur = dst;
        }
        // ---- decoder blocks
        for (int j = 0; j < S.n_up; ++j) {
            const CaeStemUp& L = S.up[j];
            const int C = L.Cout, HW = L.Hout * L.Wout, yp = C * HW, cp = 2 * yp;
            float* y = sm + m.y;
            float* cat = sm + m.cat;
            const float* w = weights(m.w_up[j], L.w, 1, L.Cout, L.Cin, L.k * L.k);
            const bool blk = L.Cout % STEM_COB == 0, s2 = L.stride == 2 && L.k <= 4;
            if (s2 && blk) stem_up<true, STEM_COB>(L, w, cur, y);
            else if (s2) stem_up<true, 1>(L, w, cur, y);
            else stem_up<false, 1>(L, w, cur, y);
            __syncthreads();
            float* avg = sm + m.small;                     // [ST][C]
            float* mx = avg + STEM_ST * C;                 // [ST][C]
            float* hid = mx + STEM_ST * C;                 // [ST][2][Cr]
            float* att = hid + STEM_ST * 2 * L.Cr;         // [ST][C]
            for (int e = tid; e < STEM_ST * C; e += STEM_NT) {
                const int s = e % STEM_ST, c = e / STEM_ST;
                const float* q = y + stem_idx(s, c * HW);
                float sum = 0.f, mxx = -INFINITY;
                for (int i = 0; i < HW; ++i) { const float t = q[i * STEM_ST]; sum += t; mxx = fmaxf(mxx, t); }
                avg[s * C + c] = sum / (float)HW;
                mx[s * C + c] = mxx;
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
                float v = c2 < C ? att[s * C + c2] * y[stem_idx(s, r)] : skip[stem_idx(s, r - yp)];
                v = stem_bn_relu(v, L.scale, L.shift, c2);
                cat[stem_idx(s, r)] = v;
                if (last && n0 + s < N) {
                    const int yy = i / L.Wout, xx = i - yy * L.Wout;
                    a.out.p[(long long)(n0 + s) * a.out.sN + (long long)c2 * a.out.sC + (long long)yy * a.out.ld + xx] = v;
                }
            }
            __syncthreads();
            cur = cat;
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
