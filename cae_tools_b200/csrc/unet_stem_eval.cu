// Eval-mode "stem" of the UNET (everything before the last transposed convolution) as ONE kernel:
//   encoder  [Conv2d + BatchNorm2d(eval) + ReLU] x n_conv          (unet.py:73-112)
//   fc stack Linear - BatchNorm1d(eval) - ReLU - Linear - ReLU, twice  (unet.py:92-100,121-129)
//   decoder  [ConvTranspose2d, ChannelAttention gate, concat with the encoder skip, BatchNorm2d(eval), ReLU] x n_up
//            (unet.py:23-39,131-163)
// In eval mode BatchNorm is a per-channel affine, so nothing couples the samples of a batch: a CTA takes STEM_S samples
// through the whole stem with every activation in shared memory (<= 3 K floats per sample for the shipped spec); each
// layer's weights are staged through shared memory as well (every weight load is shared by the STEM_S samples) and only
// the final activated concat tensor is written to global memory.  Replaces the 11 launches before the head.
// Measured (B200): 150 us against 218 us for the chain at batch 1024, 577 against 537 us at 4096 - the inner loops are
// plain per-thread loops with run-time geometry (~1 TFLOP/s); the engine uses this kernel for batches <= 2048.
#include "capi_host.h"

#define STEM_S 4
#define STEM_NT 256

struct StemSmem {            // offsets in floats, per CTA (all STEM_S samples)
    int in0, enc[CAE_STEM_MAX], va, vb, y, cat, small, wbuf;
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
    for (int e = threadIdx.x; e < total; e += STEM_NT) {
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

// v_out[s][o] = act((W v_in[s] + b) * scale + shift)
__device__ __forceinline__ void stem_fc(const CaeStemFc& L, const float* wsm, const float* in, float* outp, int in_pitch,
                                        int out_pitch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (L.in >= 64) {                                   // long rows: one warp per output, lanes over k
        for (int o = warp; o < L.out; o += STEM_NT / 32) {
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
        for (int o = threadIdx.x; o < L.out; o += STEM_NT) {
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
    for (int e = threadIdx.x; e < total; e += STEM_NT) {
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
    // layer weights go through shared memory: with ~190 KB of the SM carved out for activations the L1 keeps almost
    // nothing, and weight loads that miss it cost an L2 round trip per tap (first version: 747 us per 4096 samples)
    auto stage_w = [&](const float* w, int n) {
        for (int e = tid; e < n; e += STEM_NT) wsm[e] = __ldg(w + e);
    };
    for (int n0 = blockIdx.x * STEM_S; n0 < N; n0 += gridDim.x * STEM_S) {
        __syncthreads();
        // ---- input (samples beyond N are zero-filled; their results are never written)
        for (int e = tid; e < STEM_S * in_elems; e += STEM_NT) {
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
            stage_w(L.w, L.Cout * L.Cin * L.k * L.k);
            __syncthreads();
            stem_conv(L, wsm, cur, sm + m.enc[l], cur_pitch, op);
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
            stage_w(L.w, L.in * L.out);
            __syncthreads();
            stem_fc(L, wsm, cur, dst, cur_pitch, L.out);
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
            stage_w(L.w, L.Cin * L.Cout * L.k * L.k);
            __syncthreads();
            stem_up(L, wsm, cur, y, cur_pitch, yp);
            __syncthreads();
            float* avg = sm + m.small;                     // [S][C]
            float* mx = avg + STEM_S * C;                  // [S][C]
            float* hid = mx + STEM_S * C;                  // [S][2][Cr]
            float* att = hid + STEM_S * 2 * L.Cr;          // [S][C]
            for (int e = tid; e < STEM_S * C; e += STEM_NT) {
                const int s = e / C, c = e - s * C;
                const float* q = y + s * yp + c * HW;
                float sum = 0.f, mxx = -INFINITY;
                for (int i = 0; i < HW; ++i) { sum += q[i]; mxx = fmaxf(mxx, q[i]); }
                avg[e] = sum / (float)HW;
                mx[e] = mxx;
            }
            __syncthreads();
            for (int e = tid; e < STEM_S * 2 * L.Cr; e += STEM_NT) {
                const int s = e / (2 * L.Cr), r2 = e - s * 2 * L.Cr, which = r2 / L.Cr, r = r2 - which * L.Cr;
                const float* src = (which ? mx : avg) + s * C;
                float v = 0.f;
                for (int c = 0; c < C; ++c) v = fmaf(__ldg(L.W1 + r * C + c), src[c], v);
                hid[e] = fmaxf(v, 0.f);
            }
            __syncthreads();
            for (int e = tid; e < STEM_S * C; e += STEM_NT) {
                const int s = e / C, c = e - s * C;
                const float* h = hid + s * 2 * L.Cr;
                float v = 0.f;
                for (int r = 0; r < L.Cr; ++r) v = fmaf(__ldg(L.W2 + c * L.Cr + r), h[r] + h[L.Cr + r], v);
                att[e] = 1.f / (1.f + expf(-v));
            }
            __syncthreads();
            // cat = relu(bn([att * y ; skip]))
            const float* skip = sm + m.enc[L.skip];
            const bool last = j == S.n_up - 1;
            for (int e = tid; e < STEM_S * cp; e += STEM_NT) {
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
    m.in0 = take(STEM_S * s.conv[0].Cin * s.conv[0].Hin * s.conv[0].Win);
    int prev = s.conv[0].Cin * s.conv[0].Hin * s.conv[0].Win;
    for (int l = 0; l < s.n_conv; ++l) {
        const CaeStemConv& L = s.conv[l];
        if (!L.w || L.Cin * L.Hin * L.Win != prev) STEM_FAIL("encoder layer %d does not chain", l);
        if ((L.Hin + 2 * L.pad - L.k) / L.stride + 1 != L.Hout || (L.Win + 2 * L.pad - L.k) / L.stride + 1 != L.Wout)
            STEM_FAIL("encoder layer %d geometry", l);
        prev = L.Cout * L.Hout * L.Wout;
        m.enc[l] = take(STEM_S * prev);
    }
    int vmax = 0;
    for (int l = 0; l < s.n_fc; ++l) {
        if (!s.fc[l].w || s.fc[l].in != prev) STEM_FAIL("fc layer %d does not chain (%d vs %d)", l, s.fc[l].in, prev);
        prev = s.fc[l].out;
        vmax = max(vmax, prev);
    }
    m.va = take(STEM_S * vmax);
    m.vb = take(STEM_S * vmax);
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
    m.y = take(STEM_S * ymax);
    m.cat = take(STEM_S * 2 * ymax);
    m.small = take(STEM_S * smallmax);
    int wmax = 0;
    for (int l = 0; l < s.n_conv; ++l) wmax = max(wmax, s.conv[l].Cout * s.conv[l].Cin * s.conv[l].k * s.conv[l].k);
    for (int l = 0; l < s.n_fc; ++l) wmax = max(wmax, s.fc[l].in * s.fc[l].out);
    for (int j = 0; j < s.n_up; ++j) wmax = max(wmax, s.up[j].Cin * s.up[j].Cout * s.up[j].k * s.up[j].k);
    m.wbuf = take(wmax);
    m.total = off;
    if ((size_t)off * 4 > 160 * 1024) STEM_FAIL("activations of %d samples need %d KB of shared memory (> 160)", STEM_S, off * 4 / 1024);
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
        cudaFuncSetAttribute(k_unet_stem_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        opted = true;
    }
    const int passes = ceil_div(x->t0.N, STEM_S);
    const int per_sm = max(1, min(8, (int)((200 * 1024) / (smem + 1024))));
    const int grid = min(passes, CAE_NUM_SMS * per_sm);
    k_unet_stem_eval<<<grid, STEM_NT, smem, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_unet_stem_eval");
}
