// C ABI of libcae_b200.so: argument checking, kernel selection, launches.  See include/cae_b200.h.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_set>
#include "conv_family.cuh"
#include "conv_tiled.cuh"
#include "conv_direct.cuh"
#include "dense_misc.cuh"
#include "unet_ops.cuh"

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

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

static int check_view(const CaeView& v, const char* name) {
    CAE_REQUIRE(v.p != nullptr, "%s: null pointer", name);
    CAE_REQUIRE(v.N > 0 && v.C > 0 && v.H > 0 && v.W > 0, "%s: empty tensor %dx%dx%dx%d", name, v.N, v.C, v.H, v.W);
    CAE_REQUIRE(v.ld >= v.W, "%s: ld %d < W %d", name, v.ld, v.W);
    CAE_REQUIRE((long long)v.N * v.C * v.H * v.W < (1ll << 31), "%s: tensor too large for 32-bit positions", name);
    return CAE_OK;
}

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
static const int kTileSmemBudget = 72 * 1024;
static const int kTileSmemMax = 100 * 1024;
// kernel selection mask: bit 0 tiled up/down (k_up2/k_down2), bit 1 position-parallel wgrad (k_wgrad2a),
// bit 2 GEMM-like wgrad (k_wgrad2b).  cae_set_kernel_generation(1) = generic kernels only, (2) = default mask.
#define CAE_V2_UPDOWN 1
#define CAE_V2_WGRAD_A 2
#define CAE_V2_WGRAD_B 4
#define CAE_V2_UPDOWN_WIDE 8   // tiled up/down also for wide layers (register tile of 2/4 positions)
#define CAE_V3_DIRECT 16       // vectorised direct kernels for wide thin layers (k_up3 / k_down3)
static int default_mask() {
    const char* e = getenv("CAE_KERNEL_MASK");
    return e ? atoi(e) : (CAE_V2_UPDOWN | CAE_V2_WGRAD_A | CAE_V3_DIRECT);
}
static int g_mask = default_mask();
#define g_use_v2 (g_mask & CAE_V2_UPDOWN)
extern "C" void cae_set_kernel_generation(int gen) { g_mask = (gen <= 1) ? 0 : (gen == 2 ? default_mask() : (gen >> 4)); }

static inline int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static inline int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
static inline int roundup4(int v) { return (v + 3) & ~3; }

// opt in to > 48 KB dynamic shared memory, once per kernel (keyed by the function address)
template <typename K>
static void ensure_smem(K kernel) {
    static std::mutex mu;
    static std::unordered_set<const void*> seen;
    std::lock_guard<std::mutex> lock(mu);
    if (seen.insert((const void*)kernel).second)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemMax);
}

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
static bool aligned4(const void* p, int ld, long long sC, long long sN) {
    return p && ((uintptr_t)p % 16 == 0) && ld % 4 == 0 && sC % 4 == 0 && sN % 4 == 0;
}
static bool view_aligned(const CaeView& v) { return aligned4(v.p, v.ld, v.sC, v.sN); }
static bool src_aligned(const CaeSrc& s) {
    if (!view_aligned(s.t0)) return false;
    if (s.t1 && (uintptr_t)s.t1 % 16 != 0) return false;
    if (s.cursor && s.cursor_stride % 4 != 0) return false;
    return true;
}
static bool direct_ok(const ConvArgs& a, int width_in) {
    if (!(g_mask & CAE_V3_DIRECT)) return false;
    if (a.s != 2 || a.p != 0 || a.kh != a.kw || (a.kh != 3 && a.kh != 4)) return false;
    if (width_in < 24) return false;                                    // narrow layers: tiled kernels
    if (!src_aligned(a.in) || !view_aligned(a.out)) return false;
    if (a.epi.mode == CAE_EPI_MASKSTATS && !view_aligned(a.epi.act)) return false;
    if (a.epi.mode == CAE_EPI_SIGMOID_MSE && (!src_aligned(a.epi.target) || a.epi.target.t1 || a.epi.target.relu)) return false;
    return true;
}
static int direct_cot(int Cout) { return Cout >= 4 ? 4 : (Cout >= 2 ? 2 : 1); }

template <int K>
static int launch_up3(ConvArgs& a, cudaStream_t st, bool& handled) {
    handled = false;
    if (!direct_ok(a, a.in.t0.W)) return CAE_OK;
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

template <int K>
static int launch_down3(ConvArgs& a, cudaStream_t st, bool& handled) {
    handled = false;
    if (!direct_ok(a, a.out.W)) return CAE_OK;
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
    if (!v1_only_up && g_use_v2 && a.s == 2 && a.kh == a.kw && (a.kh == 3 || a.kh == 4)) {
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
    if (!v1_only_dn && g_use_v2 && a.s == 2 && a.kh == a.kw && (a.kh == 3 || a.kh == 4)) {
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
    dim3 grid(min(ceil_div(a.total, CAE_NT), CAE_MAX_GRID_X), out->C);
    k_ew_epilogue<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(a);
    return cae_check_launch("cae_ew_epilogue");
}

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
    int kind;          // 0: v1, 1: v2a (position parallel), 2: v2b (GEMM-like)
    int cst, cbt, cx;
    Wg2Plan a;
    WgGemmPlan b;
    size_t smem;
    int grid_x, grid_y;
    long long partials;
    StripPlan strip;
    int tiles_b;
};

static Wg2Choice plan_wgrad2(int N, int Cs, int Hs, int Ws, int Cb, int kh, int kw, int s, bool direct = false) {
    Wg2Choice w{};
    w.kind = 0;
    if (!(g_mask & (CAE_V2_WGRAD_A | CAE_V2_WGRAD_B | CAE_V3_DIRECT)) || s != 2 || kh != kw || (kh != 3 && kh != 4)) return w;
    const int KK = kh * kw;
    const long long nelem = (long long)Cs * Cb * KK;
    int cst = Cs >= 4 ? 4 : (Cs >= 2 ? 2 : 1);
    int cbt = Cb >= 2 ? 2 : 1;
    if (KK > 12 && cst == 4) cst = 2;
    if (cst == 4 && cbt == 1) cst = 2;
    if (cst == 1 && cbt == 2) cbt = 1;                      // instantiated: (4,2) (2,2) (2,1) (1,1)
    const int tiles_b = (Cb + cbt - 1) / cbt;
    const int G = ((Cs + cst - 1) / cst) * tiles_b;
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
    Wg2Choice w3 = plan_wgrad2(sm->t0.N, sm->t0.C, sm->t0.H, sm->t0.W, bg->t0.C, g->kh, g->kw, g->stride, true);
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
    const bool direct = g->pad == 0 && src_aligned(*sm) && src_aligned(*bg);
    Wg2Choice w2 = plan_wgrad2(sm->t0.N, a.Cs, sm->t0.H, sm->t0.W, a.Cb, a.kh, a.kw, a.s, direct);
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

// ---- dense / misc ------------------------------------------------------------------------
extern "C" int cae_gemm(const CaeGemm* g, void* stream) {
    CAE_REQUIRE(g && g->A && g->B && g->C, "gemm: null argument");
    CAE_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm: empty problem %dx%dx%d", g->M, g->N, g->K);
    CAE_REQUIRE((!g->a_k0 || (g->a_k2 && g->a_hw > 0)) && (!g->b_k0 || (g->b_k2 && g->b_hw > 0)),
                "gemm: on-load affine needs k0, k2 and hw");
    if (!g->a_k0 && !g->b_k0 && !g->a_relu && !g->b_relu && !g->rowsum_A && (long long)g->M * g->N <= 8192) {
        k_gemm_skinny<<<ceil_div((long long)g->M * g->N, CAE_NT), CAE_NT, 0, (cudaStream_t)stream>>>(*g);
        return cae_check_launch("cae_gemm(skinny)");
    }
    dim3 grid(ceil_div(g->N, GT), ceil_div(g->M, GT));
    k_gemm<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(*g);
    return cae_check_launch("cae_gemm");
}

extern "C" int cae_bn_eval_prepare(const CaeBN* device_table, int count, void* stream) {
    CAE_REQUIRE(device_table && count > 0, "bn_eval_prepare: bad argument");
    k_bn_eval_prepare<<<count, 128, 0, (cudaStream_t)stream>>>(device_table, count);
    return cae_check_launch("cae_bn_eval_prepare");
}

extern "C" int cae_mse(const float* a, const float* b, long long n, double* partials, unsigned int* ticket,
                       float* loss_out, const int* cursor, void* stream) {
    CAE_REQUIRE(a && b && partials && ticket && loss_out && n > 0, "mse: bad argument");
    int grid = min(ceil_div(n, CAE_NT * 4), CAE_MAX_GRID_X);
    k_mse<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(a, b, n, partials, ticket, loss_out, cursor);
    return cae_check_launch("cae_mse");
}

extern "C" int cae_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int decoupled, float grad_scale, const int* step_count,
                        void* stream) {
    CAE_REQUIRE(p && g && m && v && step_count && n > 0, "adam: bad argument");
    int grid = min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 8);
    k_adam<<<grid, CAE_NT, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled,
                                                       grad_scale, step_count);
    return cae_check_launch("cae_adam");
}

extern "C" int cae_step_advance(int* step_count, int* cursor, int n_batches, void* stream) {
    CAE_REQUIRE(step_count || cursor, "step_advance: nothing to do");
    k_step_advance<<<1, 32, 0, (cudaStream_t)stream>>>(step_count, cursor, n_batches);
    return cae_check_launch("cae_step_advance");
}

// ---- variational bottleneck ----------------------------------------------------------------------
extern "C" int cae_vae_reparam_fwd(const float* mu, const float* logvar, const float* eps, long long eps_stride,
                                   const int* cursor, float* z, int n_samples, int latent, int sample, float kl_scale,
                                   float* kl_out, void* stream) {
    CAE_REQUIRE(mu && logvar && z && n_samples > 0 && latent > 0, "vae_reparam_fwd: bad argument");
    CAE_REQUIRE(!sample || eps, "vae_reparam_fwd: sampling needs eps");
    k_vae_reparam_fwd<<<1, CAE_NT, 0, (cudaStream_t)stream>>>(mu, logvar, eps, eps_stride, cursor, z,
                                                               n_samples * latent, sample, kl_scale / (float)n_samples,
                                                               kl_out);
    return cae_check_launch("cae_vae_reparam_fwd");
}

extern "C" int cae_vae_reparam_bwd(const float* dz, const float* mu, const float* logvar, const float* eps,
                                   long long eps_stride, const int* cursor, float* dmu, float* dlogvar, int n_samples,
                                   int latent, float kl_weight, void* stream) {
    CAE_REQUIRE(dz && mu && logvar && eps && dmu && dlogvar && n_samples > 0 && latent > 0, "vae_reparam_bwd: bad argument");
    int n = n_samples * latent;
    k_vae_reparam_bwd<<<min(ceil_div(n, CAE_NT), CAE_NUM_SMS), CAE_NT, 0, (cudaStream_t)stream>>>(
        dz, mu, logvar, eps, eps_stride, cursor, dmu, dlogvar, n, kl_weight / (float)n_samples);
    return cae_check_launch("cae_vae_reparam_bwd");
}

extern "C" int cae_add2(const float* a, const float* b, float* out, long long n, void* stream) {
    CAE_REQUIRE(a && b && out && n > 0, "add2: bad argument");
    k_add2<<<min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 4), CAE_NT, 0, (cudaStream_t)stream>>>(a, b, out, n);
    return cae_check_launch("cae_add2");
}

extern "C" int cae_randn(float* out, long long n, unsigned long long seed, const int* step_count, void* stream) {
    CAE_REQUIRE(out && n > 0, "randn: bad argument");
    k_randn<<<min(ceil_div(n, CAE_NT), CAE_NUM_SMS * 4), CAE_NT, 0, (cudaStream_t)stream>>>(out, n, seed, step_count);
    return cae_check_launch("cae_randn");
}

// ---- UNET pieces ---------------------------------------------------------------------------------------
extern "C" int cae_plane_stats(const CaeView* y, float* stats, void* stream) {
    CAE_REQUIRE(y && stats, "plane_stats: null argument");
    int rc = check_view(*y, "plane_stats input");
    if (rc) return rc;
    k_plane_stats<<<y->N * y->C, CAE_NT, 0, (cudaStream_t)stream>>>(*y, stats);
    return cae_check_launch("cae_plane_stats");
}

extern "C" int cae_channel_attention_fwd(const float* stats, const float* W1, const float* W2, int N, int C, int Cr, int HW,
                                         float* att, float* hid, void* stream) {
    CAE_REQUIRE(stats && W1 && W2 && att && hid && N > 0 && C > 0 && Cr > 0 && HW > 0, "channel_attention_fwd: bad argument");
    size_t smem = (size_t)(2 * C + 2 * Cr) * 4;
    CAE_REQUIRE(smem <= 48 * 1024, "channel_attention_fwd: %d channels do not fit", C);
    k_ca_fwd<<<N, CAE_NT, smem, (cudaStream_t)stream>>>(stats, W1, W2, C, Cr, 1.f / (float)HW, att, hid);
    return cae_check_launch("cae_channel_attention_fwd");
}

extern "C" int cae_channel_attention_bwd(const float* datt, const float* att, const float* hid, const float* stats,
                                         const float* W1, const float* W2, int N, int C, int Cr, int HW, float* dW1,
                                         float* dW2, float* davg, float* dmax, void* stream) {
    CAE_REQUIRE(datt && att && hid && stats && W1 && W2 && dW1 && dW2 && davg && dmax && N > 0 && C > 0 && Cr > 0,
                "channel_attention_bwd: bad argument");
    size_t smem = (size_t)(3 * C + 2 * Cr) * 4;
    CAE_REQUIRE(smem <= 48 * 1024, "channel_attention_bwd: %d channels do not fit", C);
    k_ca_bwd<<<1, CAE_NT, smem, (cudaStream_t)stream>>>(datt, att, hid, stats, W1, W2, N, C, Cr, 1.f / (float)HW, dW1,
                                                          dW2, davg, dmax);
    return cae_check_launch("cae_channel_attention_bwd");
}

extern "C" int cae_plane_dot(const CaeSrc* g, const CaeView* y, float* out, void* stream) {
    CAE_REQUIRE(g && y && out, "plane_dot: null argument");
    CAE_REQUIRE(g->t0.N == y->N && g->t0.C == y->C && g->t0.H == y->H && g->t0.W == y->W, "plane_dot: geometry mismatch");
    k_plane_dot<<<y->N * y->C, CAE_NT, 0, (cudaStream_t)stream>>>(*g, *y, out);
    return cae_check_launch("cae_plane_dot");
}

extern "C" int cae_gate_bwd(const CaeSrc* g, const float* att, const float* davg, const float* dmax, const float* stats,
                            const CaeView* dy, float* plane_sum, void* stream) {
    CAE_REQUIRE(g && att && davg && dmax && stats && dy, "gate_bwd: null argument");
    CAE_REQUIRE(g->t0.N == dy->N && g->t0.C == dy->C && g->t0.H == dy->H && g->t0.W == dy->W, "gate_bwd: geometry mismatch");
    k_gate_bwd<<<dy->N * dy->C, CAE_NT, 0, (cudaStream_t)stream>>>(*g, att, davg, dmax, stats, *dy, plane_sum);
    return cae_check_launch("cae_gate_bwd");
}

extern "C" int cae_sum_over_n(const float* in, int N, int C, float* out, void* stream) {
    CAE_REQUIRE(in && out && N > 0 && C > 0, "sum_over_n: bad argument");
    k_sum_over_n<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(in, N, C, out);
    return cae_check_launch("cae_sum_over_n");
}

extern "C" int cae_masked_pearson_loss(const CaeView* pred, const CaeSrc* target, const CaeSrc* mask, int mask_channels,
                                       float lambda_pearson, float count_scale, double* moments, float* coef,
                                       float* scalars, float* loss_out, float* pearson_out, const CaeView* dz,
                                       float* plane_sum, void* stream) {
    CAE_REQUIRE(pred && target && moments && coef && scalars, "masked_pearson_loss: null argument");
    int rc = check_view(*pred, "masked_pearson_loss pred");
    if (rc) return rc;
    const CaeView& t = target->t0;
    CAE_REQUIRE(t.p && t.N == pred->N && t.C == pred->C && t.H == pred->H && t.W == pred->W,
                "masked_pearson_loss: target geometry differs from prediction");
    MaskedPearsonArgs a;
    memset(&a, 0, sizeof(a));
    a.pred = *pred;
    a.target = *target;
    if (mask && mask->t0.p) {
        a.mask = *mask;
        CAE_REQUIRE((mask_channels == 1 || mask_channels == pred->C) && mask->t0.C == mask_channels &&
                        mask->t0.H == pred->H && mask->t0.W == pred->W && mask->t0.N == pred->N,
                    "masked_pearson_loss: mask must be [N, 1 or C, H, W]");
        a.mask_channels = mask_channels;
    } else {
        a.mask_channels = pred->C;
    }
    a.moments = moments; a.coef = coef; a.scalars = scalars;
    a.loss_out = loss_out; a.pearson_out = pearson_out;
    a.lambda_pearson = lambda_pearson; a.count_scale = count_scale;
    cudaStream_t st = (cudaStream_t)stream;
    const int planes = pred->N * pred->C;
    k_mp_moments<<<planes, CAE_NT, 0, st>>>(a);
    k_mp_finalize<<<1, CAE_NT, 0, st>>>(a);
    if (dz) {
        CAE_REQUIRE(dz->p && dz->N == pred->N && dz->C == pred->C && dz->H == pred->H && dz->W == pred->W,
                    "masked_pearson_loss: dz geometry differs from prediction");
        k_mp_grad<<<planes, CAE_NT, 0, st>>>(a, *dz, plane_sum);
    }
    return cae_check_launch("cae_masked_pearson_loss");
}
