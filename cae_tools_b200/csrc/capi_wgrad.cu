// C ABI, weight-gradient entry points (cae_conv_wgrad): planning + launches.  See include/cae_b200.h.
#include "capi_host.h"
#include "conv_family.cuh"
#include "conv_tiled.cuh"
#include "conv_direct.cuh"
#include "conv_wgrad_tile.cuh"

// ---- weight gradient -------------------------------------------------------------------
#define CAE_WGRAD_MAX_CHUNKS 1024

struct WgradPlan {
    int cst, cbt, ntiles, tiles_b, nchunks, chunk;
    bool generic;
};

static WgradPlan plan_wgrad(int Cs, int Cb, int kh, int kw, int s, int total) {
    WgradPlan p;
    p.generic = !(s == 2 && kh >= 3 && kh <= 4 && kw >= 3 && kw <= 4);
    if (p.generic) {
        p.cst = p.cbt = 1;
        p.ntiles = Cs * Cb * kh * kw;
        p.tiles_b = Cb;
    } else {
        int kk = kh * kw;
        p.cst = Cs >= 4 ? 4 : (Cs >= 2 ? 2 : 1);
        p.cbt = Cb >= 2 ? 2 : 1;
        if (kk > 12 && p.cst == 4) p.cst = 2;                 // keep <= 64 accumulators
        if (p.cst == 4 && p.cbt == 1) p.cst = 2;              // instantiated: (4,2) (2,2) (2,1) (1,2) (1,1)
        p.tiles_b = (Cb + p.cbt - 1) / p.cbt;
        p.ntiles = ((Cs + p.cst - 1) / p.cst) * p.tiles_b;
    }
    // enough warps to fill the machine (148 SMs x 8 warps x 2), at least 64 positions per warp
    long long want = (2ll * CAE_NUM_SMS * CAE_NWARP + p.ntiles - 1) / p.ntiles;
    long long cap = (total + 63) / 64;
    long long n = want < cap ? want : cap;
    if (n < 1) n = 1;
    if (n > CAE_WGRAD_MAX_CHUNKS) n = CAE_WGRAD_MAX_CHUNKS;
    p.nchunks = (int)n;
    p.chunk = (total + p.nchunks - 1) / p.nchunks;
    p.nchunks = (total + p.chunk - 1) / p.chunk;
    return p;
}

// ---- v2 weight-gradient planning ----------------------------------------------------------------
struct Wg2Choice {
    int kind;          // 0: v1, 1: v2a (position parallel), 2: v2b (GEMM-like), 3: v3 direct, 4: tile resident
    WgTilePlan t;
    int cst, cbt, cx;
    Wg2Plan a;
    WgGemmPlan b;
    size_t smem;
    int grid_x, grid_y;
    long long partials;
    StripPlan strip;
    int tiles_b;
};

static Wg2Choice plan_wgrad2(int N, int Cs, int Hs, int Ws, int Cb, int kh, int kw, int s, bool direct = false,
                             bool s_t1 = false, bool b_t1 = false) {
    Wg2Choice w{};
    w.kind = 0;
    if (!(g_mask & (CAE_V2_WGRAD_A | CAE_V2_WGRAD_B | CAE_V3_DIRECT | CAE_WGRAD_TILE)) || s != 2 || kh != kw || (kh != 3 && kh != 4)) return w;
    const int KK = kh * kw;
    const long long nelem = (long long)Cs * Cb * KK;
    int cst = Cs >= 4 ? 4 : (Cs >= 2 ? 2 : 1);
    int cbt = Cb >= 2 ? 2 : 1;
    if (KK > 12 && cst == 4) cst = 2;
    if (cst == 4 && cbt == 1) cst = 2;
    if (cst == 1 && cbt == 2) cbt = 1;                      // instantiated: (4,2) (2,2) (2,1) (1,1)
    const int tiles_b = (Cb + cbt - 1) / cbt;
    const int G = ((Cs + cst - 1) / cst) * tiles_b;
    if (direct && (g_mask & CAE_WGRAD_TILE) && Ws >= 64 && (long long)N * Hs * Ws >= (1ll << 20) && Cs % 4 == 0 && Cb % 4 == 0 &&
        Cs <= 64 && Cb <= 32) {
        // register tile (4,2) for 3x3, (2,2) for 4x4; 256 threads = n_cst * n_cbt tiles x PS position lanes (PS >= 4)
        const int tcs = kh == 3 ? 4 : 2;
        const int pt = (Cs / tcs) * (Cb / 2);
        if (pt >= 1 && pt <= 64 && 256 % pt == 0) {
            WgTilePlan t{};
            t.CsP = Cs + 4; t.CbP = Cb + 2;
            t.n_cst = Cs / tcs; t.n_cbt = Cb / 2;
            t.PSH = 256 / pt / 4;
            t.BC = 2 * WGT_TW + kh - 2;
            for (int th = 8; th >= 2; th >>= 1) {
                t.TH = th; t.BR = 2 * th + kh - 2;
                t.TH_shift = ilog2(th); t.inv_BR = (65536 + t.BR - 1) / t.BR;
                t.SCH = th * WGT_TW + 8;
                t.BCH = t.BR * WGT_BCR + ((8 - (t.BR * WGT_BCR) % 32) + 32) % 32;
                t.raw_s = Cs * t.SCH; t.raw_b = Cb * t.BCH;
                t.cooked = (s_t1 ? 2 : 1) * t.raw_s + (b_t1 ? 2 : 1) * t.raw_b;
                size_t fl = (size_t)t.cooked + (size_t)th * WGT_TW * t.CsP + (size_t)t.BR * t.BC * t.CbP;
                size_t red = (size_t)64 * tcs * 2 * KK;
                if (fl < red) fl = red;
                if (fl * 4 <= (size_t)kTileSmemMax && (th * WGT_TW) % (4 * t.PSH) == 0) {
                    t.tiles_y = (Hs + th - 1) / th; t.tiles_x = (Ws + WGT_TW - 1) / WGT_TW;
                    t.ntiles = N * t.tiles_y * t.tiles_x;
                    w.kind = 4; w.cst = tcs; w.cbt = 2; w.t = t; w.smem = fl * 4;
                    w.grid_x = t.ntiles < 2 * CAE_NUM_SMS ? t.ntiles : 2 * CAE_NUM_SMS;
                    w.grid_y = 1;
                    w.partials = (long long)w.grid_x * nelem;
                    return w;
                }
            }
        }
    }
    if (direct && Ws >= 24 && G <= 16 && (g_mask & CAE_V3_DIRECT)) {
        w.kind = 3; w.cst = cst; w.cbt = cbt;
        w.strip.RP = Hs; w.strip.NS = (Ws + 3) / 4; w.strip.units = N * Hs * w.strip.NS;
        w.tiles_b = tiles_b;
        long long ctas = (w.strip.units + CAE_NT - 1) / CAE_NT;
        long long cap = (2ll * CAE_NUM_SMS + G - 1) / G;
        w.grid_x = (int)(ctas < cap ? ctas : cap);
        if (w.grid_x < 1) w.grid_x = 1;
        w.grid_y = G;
        w.partials = (long long)w.grid_x * nelem;
        return w;
    }
    if (G <= 8 && (g_mask & CAE_V2_WGRAD_A)) {
        Wg2Plan p{};
        int cx = Ws >= 96 ? 4 : (Ws >= 48 ? 2 : 1);
        const int TXC = 32 * cx;
        p.RP = Hs + 1; p.total_rows = N * p.RP;
        p.TR = 8;
        p.nrow_tiles = (p.total_rows + p.TR - 1) / p.TR;
        p.ncol_tiles = (Ws + TXC - 1) / TXC;
        p.SCPs = TXC; p.SCPb = roundup4(TXC + 2);
        p.tiles_b = tiles_b; p.G = G; p.GP = pow2ceil(G);
        size_t fl = (size_t)Cs * p.TR * p.SCPs + (size_t)Cb * (2 * p.TR + kh - 2) * 2 * p.SCPb;
        size_t need = (size_t)CAE_NWARP * cst * cbt * KK;
        if (fl < need) fl = need;
        if (fl * 4 <= (size_t)kTileSmemMax) {
            w.kind = 1; w.cst = cst; w.cbt = cbt; w.cx = cx; w.a = p; w.smem = fl * 4;
            long long tiles = (long long)p.nrow_tiles * p.ncol_tiles;
            w.grid_x = (int)(tiles < 2 * CAE_NUM_SMS ? tiles : 2 * CAE_NUM_SMS);
            w.grid_y = 1;
            w.partials = (long long)w.grid_x * nelem;
            return w;
        }
    }
    if (nelem >= 1024 && (g_mask & CAE_V2_WGRAD_B)) {
        WgGemmPlan p{};
        p.KK = KK; p.KW = kw;
        p.n_mtiles = (Cs + WG_BM - 1) / WG_BM;
        p.n_ntiles = (Cb * KK + WG_BN - 1) / WG_BN;
        const long long total = (long long)N * Hs * Ws;
        long long want = (2ll * CAE_NUM_SMS + p.n_mtiles * p.n_ntiles - 1) / (p.n_mtiles * p.n_ntiles);
        long long maxch = (total + WG_KS - 1) / WG_KS;
        if (want > maxch) want = maxch;
        if (want > 128) want = 128;
        if (want < 1) want = 1;
        long long kchunk = (total + want - 1) / want;
        kchunk = (kchunk + WG_KS - 1) / WG_KS * WG_KS;
        p.kchunk = (int)kchunk;
        p.nchunks = (int)((total + kchunk - 1) / kchunk);
        w.kind = 2; w.b = p; w.smem = 0;
        w.grid_x = p.nchunks; w.grid_y = p.n_mtiles * p.n_ntiles;
        w.partials = (long long)p.nchunks * nelem;
        return w;
    }
    return w;
}

template <int K, int CX>
static void launch_wgrad2a_t(const WgradArgs& a, const Wg2Choice& w, cudaStream_t st) {
    dim3 grid(w.grid_x);
    if (w.cst == 4 && w.cbt == 2) {
        if constexpr (K == 3) { ensure_smem(k_wgrad2a<K, K, CX, 4, 2>); k_wgrad2a<K, K, CX, 4, 2><<<grid, CAE_NT, w.smem, st>>>(a, w.a); }
    } else if (w.cst == 2 && w.cbt == 2) {
        ensure_smem(k_wgrad2a<K, K, CX, 2, 2>); k_wgrad2a<K, K, CX, 2, 2><<<grid, CAE_NT, w.smem, st>>>(a, w.a);
    } else if (w.cst == 2 && w.cbt == 1) {
        ensure_smem(k_wgrad2a<K, K, CX, 2, 1>); k_wgrad2a<K, K, CX, 2, 1><<<grid, CAE_NT, w.smem, st>>>(a, w.a);
    } else {
        ensure_smem(k_wgrad2a<K, K, CX, 1, 1>); k_wgrad2a<K, K, CX, 1, 1><<<grid, CAE_NT, w.smem, st>>>(a, w.a);
    }
}

template <int K>
static void launch_wgrad2a(const WgradArgs& a, const Wg2Choice& w, cudaStream_t st) {
    if (w.cx == 4) launch_wgrad2a_t<K, 4>(a, w, st);
    else if (w.cx == 2) launch_wgrad2a_t<K, 2>(a, w, st);
    else launch_wgrad2a_t<K, 1>(a, w, st);
}

static int check_wgrad(const CaeSrc* sm, const CaeSrc* bg, const CaeConvGeom* g) {
    CAE_REQUIRE(sm && bg && g, "wgrad: null argument");
    int rc;
    if ((rc = check_view(sm->t0, "wgrad small operand"))) return rc;
    if ((rc = check_view(bg->t0, "wgrad big operand"))) return rc;
    CAE_REQUIRE(sm->t0.N == bg->t0.N, "wgrad: batch mismatch");
    CAE_REQUIRE(g->kh > 0 && g->kw > 0 && g->stride > 0 && g->pad >= 0, "wgrad: bad geometry");
    return CAE_OK;
}

extern "C" long long cae_wgrad_partials_len(const CaeSrc* sm, const CaeSrc* bg, const CaeConvGeom* g) {
    if (check_wgrad(sm, bg, g)) return -1;
    WgradPlan p = plan_wgrad(sm->t0.C, bg->t0.C, g->kh, g->kw, g->stride, sm->t0.N * sm->t0.H * sm->t0.W);
    long long v1 = (long long)p.nchunks * sm->t0.C * bg->t0.C * g->kh * g->kw;
    Wg2Choice w = plan_wgrad2(sm->t0.N, sm->t0.C, sm->t0.H, sm->t0.W, bg->t0.C, g->kh, g->kw, g->stride);
    Wg2Choice w3 = plan_wgrad2(sm->t0.N, sm->t0.C, sm->t0.H, sm->t0.W, bg->t0.C, g->kh, g->kw, g->stride, true, sm->t1 != nullptr,
                               bg->t1 != nullptr);
    long long need = v1;
    if (w.kind != 0 && w.partials > need) need = w.partials;
    if (w3.kind != 0 && w3.partials > need) need = w3.partials;
    return need;    // any generation may be selected at run time
}

template <int KH, int KW, int S>
static int launch_wgrad_t(WgradArgs& a, const WgradPlan& p, cudaStream_t st) {
    int grid = ceil_div((long long)p.ntiles * p.nchunks, CAE_NWARP);
    constexpr bool big = KH * KW > 12;
    if (p.cst == 4 && p.cbt == 2) {
        if constexpr (!big) k_conv_wgrad<KH, KW, S, 4, 2><<<grid, CAE_NT, 0, st>>>(a);
    } else if (p.cst == 2 && p.cbt == 2) k_conv_wgrad<KH, KW, S, 2, 2><<<grid, CAE_NT, 0, st>>>(a);
    else if (p.cst == 2 && p.cbt == 1) k_conv_wgrad<KH, KW, S, 2, 1><<<grid, CAE_NT, 0, st>>>(a);
    else if (p.cst == 1 && p.cbt == 2) k_conv_wgrad<KH, KW, S, 1, 2><<<grid, CAE_NT, 0, st>>>(a);
    else k_conv_wgrad<KH, KW, S, 1, 1><<<grid, CAE_NT, 0, st>>>(a);
    return cae_check_launch("cae_conv_wgrad");
}

extern "C" int cae_conv_wgrad(const CaeSrc* sm, const CaeSrc* bg, const CaeConvGeom* g, float* grad, float* partials,
                              unsigned int* ticket, void* stream) {
    int rc = check_wgrad(sm, bg, g);
    if (rc) return rc;
    CAE_REQUIRE(grad && partials && ticket, "wgrad: null output/workspace");
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.sm = *sm; a.bg = *bg;
    a.kh = g->kh; a.kw = g->kw; a.s = g->stride; a.p = g->pad;
    a.grad = grad; a.partials = partials; a.ticket = ticket;
    a.Cs = sm->t0.C; a.Cb = bg->t0.C;
    a.total = sm->t0.N * sm->t0.H * sm->t0.W;
    cudaStream_t st = (cudaStream_t)stream;
    {
        // few positions, many weight elements: warp-per-element kernel without any reduction machinery
        const long long nelem = (long long)a.Cs * a.Cb * a.kh * a.kw;
        // (measured, batch 64: 256 positions 33 -> 14 us and 19 -> 8 us; at 1024 positions a draw; at 3136 it loses)
        if ((g_mask & CAE_WGRAD_SMALL) && a.total <= 640 && nelem >= 512 && sm->kn == nullptr && bg->kn == nullptr) {
            k_wgrad_small<<<ceil_div(nelem, CAE_NWARP), CAE_NT, 0, st>>>(a);
            return cae_check_launch("cae_conv_wgrad(small)");
        }
    }
    const bool direct = g->pad == 0 && src_aligned(*sm) && src_aligned(*bg);
    Wg2Choice w2 = plan_wgrad2(sm->t0.N, a.Cs, sm->t0.H, sm->t0.W, a.Cb, a.kh, a.kw, a.s, direct, sm->t1 != nullptr, bg->t1 != nullptr);
    if (w2.kind == 4) {
        if (a.kh == 3) {
            ensure_smem(k_wgrad_tile<3, 4, 2>);
            k_wgrad_tile<3, 4, 2><<<w2.grid_x, CAE_NT, w2.smem, st>>>(a, w2.t);
        } else {
            ensure_smem(k_wgrad_tile<4, 2, 2>);
            k_wgrad_tile<4, 2, 2><<<w2.grid_x, CAE_NT, w2.smem, st>>>(a, w2.t);
        }
        return cae_check_launch("cae_conv_wgrad(tile)");
    }
    if (w2.kind == 3) {
        dim3 grid(w2.grid_x, w2.grid_y);
#define CAE_WG3(KK_, S_, B_) k_wgrad3<KK_, S_, B_><<<grid, CAE_NT, 0, st>>>(a, w2.strip, w2.tiles_b)
        if (a.kh == 3) {
            if (w2.cst == 4 && w2.cbt == 2) CAE_WG3(3, 4, 2);
            else if (w2.cst == 2 && w2.cbt == 2) CAE_WG3(3, 2, 2);
            else if (w2.cst == 2 && w2.cbt == 1) CAE_WG3(3, 2, 1);
            else CAE_WG3(3, 1, 1);
        } else {
            if (w2.cst == 2 && w2.cbt == 2) CAE_WG3(4, 2, 2);
            else if (w2.cst == 2 && w2.cbt == 1) CAE_WG3(4, 2, 1);
            else CAE_WG3(4, 1, 1);
        }
#undef CAE_WG3
        return cae_check_launch("cae_conv_wgrad(v3)");
    }
    if (w2.kind == 1) {
        if (a.kh == 3) launch_wgrad2a<3>(a, w2, st); else launch_wgrad2a<4>(a, w2, st);
        return cae_check_launch("cae_conv_wgrad(v2a)");
    }
    if (w2.kind == 2) {
        k_wgrad2b<<<dim3(w2.grid_x, w2.grid_y), CAE_NT, 0, st>>>(a, w2.b);
        return cae_check_launch("cae_conv_wgrad(v2b)");
    }
    WgradPlan p = plan_wgrad(a.Cs, a.Cb, a.kh, a.kw, a.s, a.total);
    a.chunk = p.chunk; a.tiles_b = p.tiles_b; a.ntiles = p.ntiles; a.nchunks = p.nchunks;
    if (!p.generic) {
        if (a.kh == 3 && a.kw == 3) return launch_wgrad_t<3, 3, 2>(a, p, st);
        if (a.kh == 4 && a.kw == 4) return launch_wgrad_t<4, 4, 2>(a, p, st);
        if (a.kh == 4 && a.kw == 3) return launch_wgrad_t<4, 3, 2>(a, p, st);
        return launch_wgrad_t<3, 4, 2>(a, p, st);
    }
    int grid = ceil_div((long long)p.ntiles * p.nchunks, CAE_NWARP);
    k_conv_wgrad_generic<<<grid, CAE_NT, 0, st>>>(a);
    return cae_check_launch("cae_conv_wgrad(generic)");
}
