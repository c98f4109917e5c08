// C ABI of libcae_b200.so: argument checking, kernel selection, launches.  See include/cae_b200.h.
#include "capi_host.h"
#include "conv_family.cuh"
#include "conv_tiled.cuh"
#include "conv_direct.cuh"
#include "conv_down_tile.cuh"
#include "conv_up_tile.cuh"

static thread_local char g_err[512] = "";

void cae_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cae_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cae_set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return CAE_OK;
}

extern "C" const char* cae_last_error(void) { return g_err; }
extern "C" int cae_version(void) { return 100; }
extern "C" long long cae_partials_len(int C) { return (long long)CAE_MAX_GRID_X * C * 2; }

static int check_epilogue(const CaeEpilogue& e, const CaeView& out) {
    CAE_REQUIRE(e.mode >= CAE_EPI_PLAIN && e.mode <= CAE_EPI_MASK, "epilogue: bad mode %d", e.mode);
    if (e.mode == CAE_EPI_MASK) {
        CAE_REQUIRE(e.act.p && e.act.N == out.N && e.act.C == out.C && e.act.H == out.H && e.act.W == out.W,
                    "epilogue MASK: act view missing or of a different geometry");
    }
    if (e.addend.t0.p) {
        const CaeView& av = e.addend.t0;
        CAE_REQUIRE(av.N == out.N && av.C == out.C && av.H == out.H && av.W == out.W,
                    "epilogue: addend geometry differs from output");
    }
    if (epi_reduces(e.mode)) {
        CAE_REQUIRE(e.partials && e.ticket, "epilogue: reducing mode needs partials + ticket");
    }
    if (e.mode == CAE_EPI_STATS) {
        CAE_REQUIRE(e.bn.C == out.C && e.bn.scale && e.bn.shift && e.bn.mean && e.bn.invstd,
                    "epilogue STATS: BN block incomplete (C=%d vs out C=%d)", e.bn.C, out.C);
    }
    if (e.mode == CAE_EPI_MASKSTATS) {
        CAE_REQUIRE(e.act.p, "epilogue MASKSTATS: act view missing");
        CAE_REQUIRE(e.act.N == out.N && e.act.C == out.C && e.act.H == out.H && e.act.W == out.W,
                    "epilogue MASKSTATS: act geometry differs from output");
        CAE_REQUIRE(e.bn.C == out.C && e.bn.scale && e.bn.shift && e.bn.mean && e.bn.invstd && e.bn.bwdA &&
                        e.bn.bwdB && e.bn.bwdC,
                    "epilogue MASKSTATS: BN block incomplete");
    }
    if (e.mode == CAE_EPI_SIGMOID_MSE) {
        const CaeView& t = e.target.t0;
        CAE_REQUIRE(t.p && t.N == out.N && t.C == out.C && t.H == out.H && t.W == out.W,
                    "epilogue MSE: target geometry differs from output");
        CAE_REQUIRE(e.write_mode >= 0 && e.write_mode <= 2, "epilogue MSE: bad write_mode");
    }
    return CAE_OK;
}

// choose output channels per thread: largest of {8,4,2,1} that still yields >= 2 CTAs per SM,
// otherwise the one that maximises the CTA count
static int pick_cot(int Cout, int grid_x, int max_cot) {
    const int opts[4] = {8, 4, 2, 1};
    for (int i = 0; i < 4; ++i) {
        int c = opts[i];
        if (c > max_cot || c > Cout) continue;
        long long ctas = (long long)grid_x * ((Cout + c - 1) / c);
        if (ctas >= 2 * CAE_NUM_SMS) return c;
    }
    return 1;
}

static const int kSmemBudget = 40 * 1024;

template <int KH, int KW, int S>
static int launch_up_t(ConvArgs& a, cudaStream_t st) {
    int gx = min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X);
    int cot = pick_cot(a.Cout, gx, S * S * 8 <= 32 ? 8 : 4);
    int kk = KH * KW;
    a.ci_chunk = min(a.Cin, max(1, kSmemBudget / (kk * cot * 4)));
    size_t smem = (size_t)a.ci_chunk * kk * cot * 4;
    dim3 grid(gx, ceil_div(a.Cout, cot));
    switch (cot) {
        case 8: k_conv_up<KH, KW, S, 8><<<grid, CAE_NT, smem, st>>>(a); break;
        case 4: k_conv_up<KH, KW, S, 4><<<grid, CAE_NT, smem, st>>>(a); break;
        case 2: k_conv_up<KH, KW, S, 2><<<grid, CAE_NT, smem, st>>>(a); break;
        default: k_conv_up<KH, KW, S, 1><<<grid, CAE_NT, smem, st>>>(a); break;
    }
    return cae_check_launch("cae_conv_up");
}

template <int KH, int KW, int S>
static int launch_down_t(ConvArgs& a, cudaStream_t st) {
    int gx = min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X);
    int cot = pick_cot(a.Cout, gx, 8);
    int kk = KH * KW;
    a.ci_chunk = min(a.Cin, max(1, kSmemBudget / (kk * cot * 4)));
    size_t smem = (size_t)a.ci_chunk * kk * cot * 4;
    dim3 grid(gx, ceil_div(a.Cout, cot));
    switch (cot) {
        case 8: k_conv_down<KH, KW, S, 8><<<grid, CAE_NT, smem, st>>>(a); break;
        case 4: k_conv_down<KH, KW, S, 4><<<grid, CAE_NT, smem, st>>>(a); break;
        case 2: k_conv_down<KH, KW, S, 2><<<grid, CAE_NT, smem, st>>>(a); break;
        default: k_conv_down<KH, KW, S, 1><<<grid, CAE_NT, smem, st>>>(a); break;
    }
    return cae_check_launch("cae_conv_down");
}

static int fill_conv_args(ConvArgs& a, const CaeSrc* in, const float* weight, const CaeConvGeom* g, const CaeView* out,
                          const CaeEpilogue* epi) {
    CAE_REQUIRE(in && weight && g && out && epi, "conv: null argument");
    int rc;
    if ((rc = check_view(in->t0, "conv input"))) return rc;
    if ((rc = check_view(*out, "conv output"))) return rc;
    CAE_REQUIRE(g->kh > 0 && g->kw > 0 && g->stride > 0 && g->pad >= 0, "conv: bad geometry k=%dx%d s=%d p=%d", g->kh,
                g->kw, g->stride, g->pad);
    CAE_REQUIRE(in->t0.N == out->N, "conv: batch mismatch %d vs %d", in->t0.N, out->N);
    CAE_REQUIRE(in->kn == nullptr, "conv: per-(n,c) multipliers (kn) are only supported by cae_ew_epilogue");
    memset(&a, 0, sizeof(a));
    a.in = *in;
    a.w = weight;
    a.kh = g->kh; a.kw = g->kw; a.s = g->stride; a.p = g->pad;
    a.out = *out;
    a.epi = *epi;
    if (a.epi.mode == CAE_EPI_MASKSTATS && a.epi.act.p == nullptr) a.epi.mode = CAE_EPI_PLAIN;
    if ((rc = check_epilogue(a.epi, *out))) return rc;
    a.Cin = in->t0.C;
    a.Cout = out->C;
    a.inv_count = (float)((a.epi.count_scale > 0.f ? (double)a.epi.count_scale : 1.0) /
                          ((double)out->N * out->C * out->H * out->W));
    return CAE_OK;
}


// =====================================================================================================
// v2 (tiled) dispatch: stride 2, square 3x3 / 4x4 kernels.  Anything else keeps the v1 kernels.
// =====================================================================================================
static int default_mask() {
    const char* e = getenv("CAE_KERNEL_MASK");
    return e ? atoi(e) : (CAE_V2_UPDOWN | CAE_V2_WGRAD_A | CAE_V3_DIRECT | CAE_WGRAD_TILE | CAE_DOWN_TILE | CAE_UP_TILE);
}
int g_cae_mask = default_mask();
extern "C" void cae_set_kernel_generation(int gen) { g_mask = (gen <= 1) ? 0 : (gen == 2 ? default_mask() : (gen >> 4)); }

struct TileChoice { int cot, cx, nt; TilePlan plan; size_t smem; bool ok; };

// Pick the thread layout of a tiled kernel: 256 threads = TXT (x, CX positions each) * TYT (rows) * TZ (groups of
// `cot` output channels) * TK (slices of the input-channel reduction).
//   rows_total : flattened rows (all samples), width: positions per row
//   staged rows per channel = row_mult * TYT + halo_rows; staged columns = col_mult * roundup4(TXT*CX + halo_cols)
static TileChoice choose_tile(int Cin, int Cout, int rows_total, int width, int halo_rows, int row_mult, int halo_cols,
                              int col_mult, int KK, int rp, int max_cot, int acc_per_chan_pos) {
    TileChoice t{};
    t.ok = false;
    const int NT = CAE_NT;
    const int CX = width >= 64 ? 4 : (width >= 32 ? 2 : 1);
    int TXT = pow2ceil((width + CX - 1) / CX);
    if (TXT > 32) TXT = 32;
    int cot = 1;
    while (cot * 2 <= max_cot && cot * 2 <= Cout) cot *= 2;
    const int groups = (Cout + cot - 1) / cot;
    const long long base = (long long)rows_total * TXT * ((width + TXT * CX - 1) / (TXT * CX)) * groups;
    // split the reduction when the layer alone cannot fill the machine (only with small register tiles)
    int TK = 1;
    if (CX <= 2) {
        while (TK < 8 && TK * 2 <= Cin && base * TK < 2ll * CAE_NUM_SMS * NT) TK *= 2;
    }
    int rest = NT / (TXT * TK);
    if (rest < 1) { TK = NT / TXT; rest = 1; }
    // channel groups in the CTA: as many as fit while keeping >= 4 rows per tile
    int TZ = 1;
    while (TZ * 2 <= pow2ceil(groups) && rest / (TZ * 2) >= 4) TZ *= 2;
    int TYT = rest / TZ;
    if (TYT < 1) TYT = 1;
    TilePlan p{};
    p.TXT = TXT; p.txt_shift = ilog2(TXT); p.RP = rp; p.total_rows = rows_total;
    p.TZ = TZ; p.tz_shift = ilog2(TZ); p.tyt_shift = ilog2(TYT); p.TK = TK;
    p.nrow_tiles = (rows_total + TYT - 1) / TYT;
    p.ncol_tiles = (width + TXT * CX - 1) / (TXT * CX);
    p.SROWS = row_mult * TYT + halo_rows;
    p.SCP = roundup4(TXT * CX + halo_cols);
    const size_t per_ch = (size_t)p.SROWS * p.SCP * col_mult * 4 + (size_t)KK * cot * TZ * 4;
    const size_t fixed = (size_t)p.SROWS * 8 + 32;
    if (per_ch + fixed > (size_t)kTileSmemMax) return t;
    int chunk = (int)((kTileSmemBudget - fixed) / per_ch);
    if (chunk < 1) chunk = 1;
    if (chunk > Cin) chunk = Cin;
    p.ci_chunk = chunk;
    size_t smem = per_ch * chunk + fixed;
    size_t scratch = (size_t)NT * 2 * cot * 4;                                   // statistics tail
    if (TK > 1) scratch = max(scratch, (size_t)NT * acc_per_chan_pos * cot * CX * 4);   // split-K partials
    if (scratch > (size_t)kTileSmemMax) return t;
    if (smem < scratch) smem = scratch;
    t.cot = cot; t.cx = CX; t.nt = NT; t.plan = p; t.smem = smem; t.ok = true;
    return t;
}

#define CAE_LAUNCH_TILED(KERNEL, KH, KW)                                                                              \
    do {                                                                                                              \
        const int key = tc.cx * 16 + tc.cot;                                                                          \
        switch (key) {                                                                                                \
            case 1 * 16 + 1: ensure_smem(KERNEL<KH, KW, 1, 1>); KERNEL<KH, KW, 1, 1><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 1 * 16 + 2: ensure_smem(KERNEL<KH, KW, 1, 2>); KERNEL<KH, KW, 1, 2><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 1 * 16 + 4: ensure_smem(KERNEL<KH, KW, 1, 4>); KERNEL<KH, KW, 1, 4><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 2 * 16 + 1: ensure_smem(KERNEL<KH, KW, 2, 1>); KERNEL<KH, KW, 2, 1><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 2 * 16 + 2: ensure_smem(KERNEL<KH, KW, 2, 2>); KERNEL<KH, KW, 2, 2><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 2 * 16 + 4: ensure_smem(KERNEL<KH, KW, 2, 4>); KERNEL<KH, KW, 2, 4><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 4 * 16 + 1: ensure_smem(KERNEL<KH, KW, 4, 1>); KERNEL<KH, KW, 4, 1><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            case 4 * 16 + 2: ensure_smem(KERNEL<KH, KW, 4, 2>); KERNEL<KH, KW, 4, 2><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
            default:         ensure_smem(KERNEL<KH, KW, 4, 4>); KERNEL<KH, KW, 4, 4><<<grid, tc.nt, tc.smem, st>>>(a, tc.plan); break; \
        }                                                                                                             \
    } while (0)

template <int KH, int KW>
static int launch_up2(ConvArgs& a, cudaStream_t st, bool& handled) {
    constexpr int JY = (KH + 1) / 2, JX = (KW + 1) / 2;
    handled = false;
    const int QH = (a.out.H - 1 + a.p) / 2 + 1, QW = (a.out.W - 1 + a.p) / 2 + 1;
    if (QH - a.in.t0.H < JY - 1) return CAE_OK;
    TileChoice tc = choose_tile(a.Cin, a.Cout, a.out.N * QH, QW, JY - 1, 1, JX - 1, 1, KH * KW, QH, 4, 4);
    if (!tc.ok || (tc.cx > 1 && !(g_mask & CAE_V2_UPDOWN_WIDE))) return CAE_OK;
    handled = true;
    a.QH = QH; a.QW = QW;
    int ntiles = tc.plan.nrow_tiles * tc.plan.ncol_tiles;
    dim3 grid(min(ntiles, CAE_MAX_GRID_X), ceil_div(a.Cout, tc.cot * tc.plan.TZ));
    CAE_LAUNCH_TILED(k_up2, KH, KW);
    return cae_check_launch("cae_conv_up(v2)");
}

template <int KH, int KW>
static int launch_down2(ConvArgs& a, cudaStream_t st, bool& handled) {
    handled = false;
    const int OHp = a.out.H + 1;
    TileChoice tc = choose_tile(a.Cin, a.Cout, a.out.N * OHp, a.out.W, KH - 2, 2, 2, 2, KH * KW, OHp, 4, 1);
    if (!tc.ok || (tc.cx > 1 && !(g_mask & CAE_V2_UPDOWN_WIDE))) return CAE_OK;
    handled = true;
    int ntiles = tc.plan.nrow_tiles * tc.plan.ncol_tiles;
    dim3 grid(min(ntiles, CAE_MAX_GRID_X), ceil_div(a.Cout, tc.cot * tc.plan.TZ));
    CAE_LAUNCH_TILED(k_down2, KH, KW);
    return cae_check_launch("cae_conv_down(v2)");
}


// ---- v3 direct kernels: eligibility + launch -------------------------------------------------------
static bool direct_ok(const ConvArgs& a, int width_in) {
    if (!(g_mask & CAE_V3_DIRECT)) return false;
    if (a.s != 2 || a.p != 0 || a.kh != a.kw || (a.kh != 3 && a.kh != 4)) return false;
    if (width_in < 24) return false;                                    // narrow layers: tiled kernels
    if (!src_aligned(a.in) || !view_aligned(a.out)) return false;
    if (a.epi.mode == CAE_EPI_MASKSTATS && !view_aligned(a.epi.act)) return false;
    if (a.epi.mode == CAE_EPI_SIGMOID_MSE && (!src_aligned(a.epi.target) || a.epi.target.t1 || a.epi.target.relu)) return false;
    return true;
}
static int direct_cot(int Cout) {
    static const int cap = [] { const char* e = getenv("CAE_DIRECT_COT"); return e ? atoi(e) : 4; }();   // tuning knob
    const int c = Cout >= 4 ? 4 : (Cout >= 2 ? 2 : 1);
    return c < cap ? c : cap;
}

template <int K, int COT>
static int launch_up_tile_t(ConvArgs& a, const UpTilePlan& p, dim3 grid, size_t smem, cudaStream_t st) {
    ensure_smem(k_up_tile<K, COT>);
    k_up_tile<K, COT><<<grid, CAE_NT, smem, st>>>(a, p);
    return cae_check_launch("cae_conv_up(tile)");
}

template <int K>
static int launch_up3(ConvArgs& a, cudaStream_t st, bool& handled) {
    handled = false;
    if (!direct_ok(a, a.in.t0.W)) return CAE_OK;
    if ((g_mask & CAE_UP_TILE) && a.in.t0.W >= 64 && a.in.t1 == nullptr && a.Cin <= 64 &&
        (long long)a.out.N * a.out.H * a.out.W >= (1ll << 21)) {
        const int cot = direct_cot(a.Cout);
        UpTilePlan p{};
        const int QH = (a.out.H - 1) / 2 + 1, QW = (a.out.W - 1) / 2 + 1;
        p.tiles_y = ceil_div(QH, UT_ROWS);
        p.tiles_x = ceil_div(QW, UT_STRIPS * 4);
        p.ntiles = a.out.N * p.tiles_y * p.tiles_x;
        p.nchunks = ceil_div(a.Cin, UT_CC);
        const size_t smem = ((size_t)a.Cin * K * K * cot + roundup4(a.Cin * 4) + 2 * (size_t)UT_CC * UT_IR * UT_IW) * 4;
        if (smem <= (size_t)kTileSmemMax) {
            const int gy = ceil_div(a.Cout, cot);
            int gx = (2 * CAE_NUM_SMS + gy - 1) / gy;
            if (gx > p.ntiles) gx = p.ntiles;
            handled = true;
            if (cot == 4) return launch_up_tile_t<K, 4>(a, p, dim3(gx, gy), smem, st);
            if (cot == 2) return launch_up_tile_t<K, 2>(a, p, dim3(gx, gy), smem, st);
            return launch_up_tile_t<K, 1>(a, p, dim3(gx, gy), smem, st);
        }
    }
    const int cot = direct_cot(a.Cout);
    const size_t smem = (size_t)a.Cin * K * K * cot * 4;
    if (smem > 48 * 1024) return CAE_OK;
    StripPlan p{};
    p.RP = (a.out.H - 1) / 2 + 1;
    const int QW = (a.out.W - 1) / 2 + 1;
    p.NS = (QW + 3) / 4;
    p.units = a.out.N * p.RP * p.NS;
    handled = true;
    dim3 grid(min(ceil_div(p.units, CAE_NT), CAE_MAX_GRID_X), ceil_div(a.Cout, cot));
    if (cot == 4) k_up3<K, 4><<<grid, CAE_NT, smem, st>>>(a, p);
    else if (cot == 2) k_up3<K, 2><<<grid, CAE_NT, smem, st>>>(a, p);
    else k_up3<K, 1><<<grid, CAE_NT, smem, st>>>(a, p);
    return cae_check_launch("cae_conv_up(v3)");
}

template <int K, int COT>
static int launch_down_tile_t(ConvArgs& a, const DownTilePlan& p, dim3 grid, size_t smem, cudaStream_t st) {
    ensure_smem(k_down_tile<K, COT>);
    k_down_tile<K, COT><<<grid, CAE_NT, smem, st>>>(a, p);
    return cae_check_launch("cae_conv_down(tile)");
}

template <int K>
static int launch_down3(ConvArgs& a, cudaStream_t st, bool& handled) {
    handled = false;
    if (!direct_ok(a, a.out.W)) return CAE_OK;
    if ((g_mask & CAE_DOWN_TILE) && a.out.W >= 64 && (long long)a.out.N * a.out.H * a.out.W >= (1ll << 19) && a.Cout % 4 == 0 &&
        a.Cin <= 64) {
        const int cot = a.Cout % 8 == 0 ? 8 : 4;
        DownTilePlan p{};
        p.tiles_y = ceil_div(a.out.H, DT_ROWS);
        p.tiles_x = ceil_div(a.out.W, DT_STRIPS * 4);
        p.ntiles = a.out.N * p.tiles_y * p.tiles_x;
        p.IR = 2 * DT_ROWS + K - 2;
        p.ntens = a.in.t1 ? 2 : 1;
        p.stage_fl = p.ntens * p.IR * DT_IW;
        const size_t smem = ((size_t)a.Cin * K * K * cot + roundup4(a.Cin * 4) + 2 * (size_t)p.stage_fl) * 4;
        if (smem <= (size_t)kTileSmemMax) {
            const int gy = a.Cout / cot;
            int gx = (2 * CAE_NUM_SMS + gy - 1) / gy;
            if (gx > p.ntiles) gx = p.ntiles;
            handled = true;
            if (cot == 8) return launch_down_tile_t<K, 8>(a, p, dim3(gx, gy), smem, st);
            return launch_down_tile_t<K, 4>(a, p, dim3(gx, gy), smem, st);
        }
    }
    const int cot = direct_cot(a.Cout);
    const size_t smem = (size_t)a.Cin * K * K * cot * 4;
    if (smem > 48 * 1024) return CAE_OK;
    StripPlan p{};
    p.RP = a.out.H;
    p.NS = (a.out.W + 3) / 4;
    p.units = a.out.N * p.RP * p.NS;
    handled = true;
    dim3 grid(min(ceil_div(p.units, CAE_NT), CAE_MAX_GRID_X), ceil_div(a.Cout, cot));
    if (cot == 4) k_down3<K, 4><<<grid, CAE_NT, smem, st>>>(a, p);
    else if (cot == 2) k_down3<K, 2><<<grid, CAE_NT, smem, st>>>(a, p);
    else k_down3<K, 1><<<grid, CAE_NT, smem, st>>>(a, p);
    return cae_check_launch("cae_conv_down(v3)");
}

extern "C" int cae_conv_up(const CaeSrc* in, const float* weight, const CaeConvGeom* g, const CaeView* out,
                           const CaeEpilogue* epi, void* stream) {
    ConvArgs a;
    int rc = fill_conv_args(a, in, weight, g, out, epi);
    if (rc) return rc;
    const CaeView& iv = in->t0;
    // Hout = (Hin-1)*s - 2p + kh + output_padding, 0 <= output_padding < s
    int hmin = (iv.H - 1) * a.s - 2 * a.p + a.kh, wmin = (iv.W - 1) * a.s - 2 * a.p + a.kw;
    CAE_REQUIRE(out->H >= hmin && out->H < hmin + max(a.s, 1) && out->W >= wmin && out->W < wmin + max(a.s, 1),
                "conv_up: output %dx%d inconsistent with input %dx%d k=%dx%d s=%d p=%d", out->H, out->W, iv.H, iv.W,
                a.kh, a.kw, a.s, a.p);
    cudaStream_t st = (cudaStream_t)stream;
    const bool v1_only_up = a.epi.addend.t0.p != nullptr || a.epi.mode == CAE_EPI_MASK;   // features of the generic epilogue
    if (!v1_only_up && a.s == 2 && a.kh == a.kw && (a.kh == 3 || a.kh == 4)) {
        bool handled = false;
        rc = (a.kh == 3) ? launch_up3<3>(a, st, handled) : launch_up3<4>(a, st, handled);
        if (handled) return rc;
    }
    // (epi_row handles addend / MASK too; measured: with an addend the tiled kernel wins from 32 input channels on -
    //  unet conv2.dgrad 23 -> 18 us - and loses below - conv1.dgrad 12.5 -> 17.8 us)
    if (g_use_v2 && a.s == 2 && a.kh == a.kw && (a.kh == 3 || a.kh == 4) && (!v1_only_up || a.Cin >= 32)) {
        bool handled = false;
        rc = (a.kh == 3) ? launch_up2<3, 3>(a, st, handled) : launch_up2<4, 4>(a, st, handled);
        if (handled) return rc;
    }
    if (a.s == 2 && a.kh >= 3 && a.kh <= 4 && a.kw >= 3 && a.kw <= 4) {
        a.QH = (out->H - 1 + a.p) / a.s + 1;
        a.QW = (out->W - 1 + a.p) / a.s + 1;
        a.total = out->N * a.QH * a.QW;
        if (a.kh == 3 && a.kw == 3) return launch_up_t<3, 3, 2>(a, st);
        if (a.kh == 4 && a.kw == 4) return launch_up_t<4, 4, 2>(a, st);
        if (a.kh == 4 && a.kw == 3) return launch_up_t<4, 3, 2>(a, st);
        return launch_up_t<3, 4, 2>(a, st);
    }
    a.QH = out->H; a.QW = out->W;
    a.total = out->N * out->H * out->W;
    dim3 grid(min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X), a.Cout);
    k_conv_up_generic<<<grid, CAE_NT, 0, st>>>(a);
    return cae_check_launch("cae_conv_up(generic)");
}

extern "C" int cae_conv_down(const CaeSrc* in, const float* weight, const CaeConvGeom* g, const CaeView* out,
                             const CaeEpilogue* epi, void* stream) {
    ConvArgs a;
    int rc = fill_conv_args(a, in, weight, g, out, epi);
    if (rc) return rc;
    const CaeView& iv = in->t0;
    // Hout <= (Hin + 2p - kh)/s + 1  (a transposed conv with output_padding leaves unused input rows)
    CAE_REQUIRE((out->H - 1) * a.s + a.kh <= iv.H + 2 * a.p && (out->W - 1) * a.s + a.kw <= iv.W + 2 * a.p,
                "conv_down: output %dx%d too large for input %dx%d k=%dx%d s=%d p=%d", out->H, out->W, iv.H, iv.W, a.kh,
                a.kw, a.s, a.p);
    cudaStream_t st = (cudaStream_t)stream;
    a.QH = out->H; a.QW = out->W;
    a.total = out->N * out->H * out->W;
    const bool v1_only_dn = a.epi.addend.t0.p != nullptr || a.epi.mode == CAE_EPI_MASK;
    if (!v1_only_dn && a.s == 2 && a.kh == a.kw && (a.kh == 3 || a.kh == 4)) {
        bool handled = false;
        rc = (a.kh == 3) ? launch_down3<3>(a, st, handled) : launch_down3<4>(a, st, handled);
        if (handled) return rc;
    }
    if (g_use_v2 && a.s == 2 && a.kh == a.kw && (a.kh == 3 || a.kh == 4)) {
        bool handled = false;
        rc = (a.kh == 3) ? launch_down2<3, 3>(a, st, handled) : launch_down2<4, 4>(a, st, handled);
        if (handled) return rc;
    }
    if (a.s == 2 && a.kh >= 3 && a.kh <= 4 && a.kw >= 3 && a.kw <= 4) {
        if (a.kh == 3 && a.kw == 3) return launch_down_t<3, 3, 2>(a, st);
        if (a.kh == 4 && a.kw == 4) return launch_down_t<4, 4, 2>(a, st);
        if (a.kh == 4 && a.kw == 3) return launch_down_t<4, 3, 2>(a, st);
        return launch_down_t<3, 4, 2>(a, st);
    }
    dim3 grid(min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X), a.Cout);
    k_conv_down_generic<<<grid, CAE_NT, 0, st>>>(a);
    return cae_check_launch("cae_conv_down(generic)");
}

extern "C" int cae_ew_epilogue(const CaeSrc* in, const CaeView* out, const CaeEpilogue* epi, void* stream) {
    CAE_REQUIRE(in && out && epi, "ew: null argument");
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    int rc;
    if ((rc = check_view(in->t0, "ew input"))) return rc;
    if ((rc = check_view(*out, "ew output"))) return rc;
    CAE_REQUIRE(in->t0.N == out->N && in->t0.C == out->C && in->t0.H == out->H && in->t0.W == out->W,
                "ew: geometry mismatch");
    a.in = *in;
    a.out = *out;
    a.epi = *epi;
    if (a.epi.mode == CAE_EPI_MASKSTATS && a.epi.act.p == nullptr) a.epi.mode = CAE_EPI_PLAIN;
    if ((rc = check_epilogue(a.epi, *out))) return rc;
    a.Cin = a.Cout = out->C;
    a.QH = out->H; a.QW = out->W;
    a.total = out->N * out->H * out->W;
    a.inv_count = (float)((a.epi.count_scale > 0.f ? (double)a.epi.count_scale : 1.0) /
                          ((double)out->N * out->C * out->H * out->W));
    if (out->H == 1 && out->W == 1 && out->C >= 64 && in->t0.sC == 1 && out->sC == 1 &&
        (a.epi.mode != CAE_EPI_MASKSTATS || a.epi.act.sC == 1) && a.epi.mode != CAE_EPI_SIGMOID_MSE &&
        (a.epi.addend.t0.p == nullptr || a.epi.addend.t0.sC == 1)) {
        k_ew_rows<<<ceil_div(out->C, 32), CAE_NT, 0, (cudaStream_t)stream>>>(a);
        return cae_check_launch("cae_ew_epilogue(rows)");
    }
    dim3 grid(min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X), out->C);
    k_ew_epilogue<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_ew_epilogue");
}
