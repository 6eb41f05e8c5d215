// C ABI of the whitomo_b200 engine (see include/whitomo_b200.h for the contract and the reference
// interfaces each entry point replaces).  Host side only: argument checks, workspace carving, launches.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "bp.cuh"
#include "common.cuh"
#include "epilogues.cuh"
#include "fp.cuh"

namespace wt {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return WT_ERR_CUDA;
}
static int invalid(const char *msg) {
    g_err = msg;
    return WT_ERR_INVALID;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carve the caller's scratch into 256-byte aligned pieces
struct Carver {
    char *base;
    size_t off, cap;
    Carver(void *p, size_t bytes) : base((char *)p), off(0), cap(bytes) {}
    template <class T> T *take(size_t count) {
        T *r = (T *)(base + off);
        off = align_up(off + count * sizeof(T), 256);
        return r;
    }
    bool ok() const { return off <= cap; }
};

static size_t vn_floats(const wt_geom *g, int n) { return (size_t)n * g->rows * fp_pitch(g->cols); }
static size_t vt_floats(const wt_geom *g, int n) { return (size_t)n * g->cols * fp_pitch(g->rows); }
static size_t packed_vol_floats(const wt_geom *g, int n) {
    return align_up(vn_floats(g, n), 64) + align_up(vt_floats(g, n), 64);
}
static size_t packed_sino_floats(const wt_geom *g, int n) { return align_up((size_t)n * g->nangles * g->sino_pitch, 64); }
static int fp_blocks(const wt_geom *g) { return ((g->ndet + FP_DT - 1) / FP_DT) * g->ntiles; }

// Images are processed in groups of V interleaved images.  Three or more images always go through the float4
// kernels, the last group padded with zero components (never stored) when the count is not a multiple of four: an
// image's bits then do not depend on how many images travel with it, and a padded float4 group is cheaper than the
// float2 + float kernels it replaces.  One or two images run as a float2 / float group, or mirror-paired (below).
struct Chunk { int v, groups; };
static Chunk chunk_of(int n) {
    if (n >= 3) return {4, (n + 3) / 4};
    return {n, 1};
}

// Calls with one or two images of a large volume use mirror pairing instead (each image also travels as its point reflection: V = 2n components, half the rays).
// Pairing halves the number of CTAs, so it only pays once the backprojector still fills the GPU with half its pixel
// tiles: images of at least 768^2 pixels (measured: 1.3-1.7x faster at 1024^2 and up, slower at 512^2 and below).
// Pairing relies on the symmetry of the exact positions c_0 + l * delta; the reference's float32 running sums are not
// symmetric (a ray and its mirror image move through different binades), so it is used under WT_POS_EXACT only.
static inline bool use_mirror(const wt_geom *g, int n) {
    return g->positions == WT_POS_EXACT && n <= 2 && (size_t)g->rows * g->cols >= (size_t)768 * 768;
}
static inline int packed_images(const wt_geom *g, int n) { return use_mirror(g, n) ? 2 * n : chunk_of(n).v * chunk_of(n).groups; }

// returns the number of per-image partial-sum slots the epilogue wrote (CTAs x passes), or a negative error
// packed_in: vn / vt already hold the packed copies of the volume (written by a previous kernel's epilogue)
template <class Epi>
static int run_fp(const wt_geom *g, const float *vol, int n, int a_begin, int a_end, float *vn, float *vt,
                  const Epi &epi_proto, cudaStream_t st, bool packed_in = false) {
    FpParams p;
    p.vn = vn; p.vt = vt; p.ang = g->d_fp;
    p.tiles = (const FpTile *)g->d_tiles; p.ntiles = g->ntiles;
    p.R = g->rows; p.C = g->cols; p.D = g->ndet; p.A = g->nangles;
    p.a_begin = a_begin; p.a_end = a_end;
    p.warpf = g->d_warpf; p.astra = g->positions == WT_POS_ASTRA;
    p.Ex = -0.5f * (float)g->cols + 0.5f; p.Ey = 0.5f * (float)g->rows - 0.5f;
    p.order = g->d_order; p.nphase = g->nphase;
    static const bool tile_major = [] { const char *e = getenv("WT_FP_ORDER"); return e && !strcmp(e, "tile"); }();
    if (tile_major) p.nphase = 0;   // A/B measurements only
    for (int i = 0; i <= FP_MAX_PHASES; ++i) p.phase_start[i] = g->phase_start[i];
    p.ndblk = 0;   // set by launch_fp from Dproc
    p.nimg = n;
    // only the packed copies the angles of this call read: an angle-range call of one orientation packs half
    int copies = 0;
    for (int a = a_begin; a < a_end; ++a) copies |= g->h_vertical[a] ? 1 : 2;
    const Epi epi = epi_proto;
    if (use_mirror(g, n)) {
        p.Dproc = (g->ndet + 1) / 2;
        if (n == 1) {
            if constexpr (Epi::kMult == 1) {
                launch_pack_vol<2, true>(vol, vn, vt, p.R, p.C, 1, n, copies, st);
                WT_CUDA((launch_fp<2, Epi, true>(p, epi, 1, st)));
            } else {
                return invalid("run_fp: image count is not a multiple of the element count");
            }
        } else {
            launch_pack_vol<4, true>(vol, vn, vt, p.R, p.C, 1, n, copies, st);
            WT_CUDA((launch_fp<4, Epi, true>(p, epi, 1, st)));
        }
        WT_CUDA(cudaGetLastError());
        return 2 * ((p.Dproc + FP_DT - 1) / FP_DT) * g->ntiles;
    }
    p.Dproc = g->ndet;
    if (n % Epi::kMult) return invalid("run_fp: image count is not a multiple of the element count");
    const Chunk c = chunk_of(n);
    switch (c.v) {
    case 4: if (!packed_in) launch_pack_vol<4>(vol, vn, vt, p.R, p.C, c.groups, n, copies, st); WT_CUDA((launch_fp<4, Epi>(p, epi, c.groups, st))); break;
    case 2:
        if constexpr (2 % Epi::kMult == 0) { if (!packed_in) launch_pack_vol<2>(vol, vn, vt, p.R, p.C, 1, n, copies, st); WT_CUDA((launch_fp<2, Epi>(p, epi, 1, st))); }
        break;
    default:
        if constexpr (Epi::kMult == 1) { if (!packed_in) launch_pack_vol<1>(vol, vn, vt, p.R, p.C, 1, n, copies, st); WT_CUDA((launch_fp<1, Epi>(p, epi, 1, st))); }
        break;
    }
    WT_CUDA(cudaGetLastError());
    return fp_blocks(g);
}

// rows [row_begin, row_end) of the rows that get a thread (all rows, or the upper half under mirror pairing, where the
// point-reflected rows are written too); row_end < 0: all of them.  packed: `sp` already holds the packed sinogram.
template <class Epi>
static int run_bp(const wt_geom *g, const float *sino, int n, int a_begin, int a_end, float *sp, const Epi &epi_proto,
                  cudaStream_t st, int row_begin = 0, int row_end = -1, bool packed = false) {
    BpParams p;
    p.sp = sp; p.bwarp = g->d_bwarp; p.bwarpd = g->d_bwarpd; p.plans = (const BpPlanRec *)g->d_bplans;
    p.R = g->rows; p.C = g->cols; p.A = g->nangles;
    p.padl = g->sino_padl; p.pitch = g->sino_pitch;
    p.a_begin = a_begin; p.a_end = a_end;
    p.nimg = n;
    const Epi epi = epi_proto;
    const bool mirror = use_mirror(g, n);
    p.Rproc = mirror ? (g->rows + 1) / 2 : g->rows;
    p.row_begin = row_begin;
    p.row_end = row_end < 0 ? p.Rproc : row_end;
    if (p.row_begin < 0 || p.row_end > p.Rproc || p.row_begin > p.row_end || p.row_begin % bp_tile_height(g->rows, g->cols))
        return invalid("run_bp: bad row range (row_begin must be a multiple of the tile height)");
    if (p.row_begin == p.row_end) return WT_OK;
    if (mirror) {
        if constexpr (!Epi::kHasSums) {
            if (n == 1) { if (!packed) launch_pack_sino<2, true>(sino, sp, g, 1, n, a_begin, a_end, st); WT_CUDA((launch_bp<2, Epi, true>(p, epi, 1, st))); }
            else { if (!packed) launch_pack_sino<4, true>(sino, sp, g, 1, n, a_begin, a_end, st); WT_CUDA((launch_bp<4, Epi, true>(p, epi, 1, st))); }
            WT_CUDA(cudaGetLastError());
            return WT_OK;
        } else {
            return invalid("run_bp: this epilogue is not mirror-paired");
        }
    }
    if (n % Epi::kMult) return invalid("run_bp: image count is not a multiple of the element count");
    const Chunk c = chunk_of(n);
    switch (c.v) {
    case 4: if (!packed) launch_pack_sino<4>(sino, sp, g, c.groups, n, a_begin, a_end, st); WT_CUDA((launch_bp<4, Epi>(p, epi, c.groups, st))); break;
    case 2:
        if constexpr (2 % Epi::kMult == 0) { if (!packed) launch_pack_sino<2>(sino, sp, g, 1, n, a_begin, a_end, st); WT_CUDA((launch_bp<2, Epi>(p, epi, 1, st))); }
        break;
    default:
        if constexpr (Epi::kMult == 1) { if (!packed) launch_pack_sino<1>(sino, sp, g, 1, n, a_begin, a_end, st); WT_CUDA((launch_bp<1, Epi>(p, epi, 1, st))); }
        break;
    }
    WT_CUDA(cudaGetLastError());
    return WT_OK;
}

// epilogues as the host instantiates them: kMult = images per problem (a group must hold whole problems)
struct HFpStore : FpEpiStore { static constexpr int kMult = 1; };
struct HFpSirt : FpEpiSirtResidual { static constexpr int kMult = 1; };
struct HFpMar : FpEpiMar { static constexpr int kMult = 1; };
static int bp_slots(const wt_geom *g) {      // partial-sum slots per slice of an epilogue with sums: (tile, consumer warp)
    const int th = bp_tile_height(g->rows, g->cols);
    return ((g->cols + BP_TW - 1) / BP_TW) * ((g->rows + th - 1) / th) * (BP_THREADS / 32);
}
template <int K> struct HFpWhite : FpEpiWhite<K> { static constexpr int kMult = K; };
typedef BpEpiStore HBpStore;
typedef BpEpiSirtUpdate HBpSirt;

// constant-memory slots of the spectrum tables, per device
static std::mutex g_slot_mutex;
static unsigned g_slots_used[64] = {};   // bit i of entry d: slot i is taken on device d

// Everything wt_geom_create derives from the geometry on the host, without touching CUDA: per-angle constants of both
// projectors, the forward projector's angle tiles and their launch order.  wt_geom_plan exposes it for CPU tests.
struct GeomPlan {
    std::vector<FpAngle> fp;
    std::vector<BpAngle> bp;
    std::vector<WarpAngle> warp;
    std::vector<WarpAngleF> warpf;
    std::vector<BpWarpF> bwarp;
    std::vector<BpWarpD> bwarpd;
    std::vector<FpTile> tiles;
    std::vector<int> order;
    int phase_start[FP_MAX_PHASES + 1];
    int nphase;
};

static int check_geometry(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, const char *who) {
    if (!angles) { g_err = std::string(who) + ": null pointer"; return WT_ERR_INVALID; }
    if (rows <= 0 || cols <= 0 || ndet <= 0 || nangles <= 0) { g_err = std::string(who) + ": sizes must be positive"; return WT_ERR_INVALID; }
    if (!(det_width >= 1.0f)) {
        g_err = std::string(who) + ": det_width < 1 needs more than two backprojection taps (unsupported)";
        return WT_ERR_UNSUPPORTED;
    }
    return WT_OK;
}

static int default_positions() {
    const char *e = getenv("WT_POSITIONS");
    return (e && !strcmp(e, "exact")) ? WT_POS_EXACT : WT_POS_ASTRA;
}

// The reference's float32 running sum as a position warp (common.cuh: WarpAngle): inside the binade [2^e, 2^(e+1)) a
// float32 addition of delta advances the position by rint(delta / u) * u, u = 2^(e-23) (round half to even is what a
// tie settles into after its first step).  Where the step rounds to zero (|delta| < u/2) the reference's ray stands
// still: G gets a huge slope there, so rays neither leave such a stripe nor travel into it beyond its edge.
static WarpAngle make_warp(float delta, int positions) {
    WarpAngle w;
    double g = 0.0;
    for (int i = 0; i < WT_NB; ++i) {
        double slope = 1.0;
        if (positions == WT_POS_ASTRA && i > 0 && delta != 0.0f) {
            const double u = ldexp(1.0, 1 + i - 23);          // stripe i = binade [2^(i+1), 2^(i+2))
            const double step = nearbyint((double)delta / u) * u;
            slope = step != 0.0 ? (double)delta / step : WT_STUCK_SLOPE;
        }
        w.gcum[i] = g;
        w.slope[i] = slope;
        w.islope[i] = 1.0 / slope;
        g += (stripe_pos(i + 1) - stripe_pos(i)) * slope;
    }
    return w;
}

static int plan_geometry(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, int positions,
                         GeomPlan &pl) {
    pl.fp.assign(nangles, FpAngle());
    pl.bp.assign(nangles, BpAngle());
    pl.warp.assign(nangles, WarpAngle());
    pl.warpf.assign(nangles, WarpAngleF());
    pl.bwarp.assign(nangles, BpWarpF());
    pl.bwarpd.assign(nangles, BpWarpD());
    std::vector<FpAngle> &fp = pl.fp;
    std::vector<BpAngle> &bp = pl.bp;
    const double Ex = -0.5 * cols + 0.5, Ey = 0.5 * rows - 0.5;
    for (int a = 0; a < nangles; ++a) {
        // the float32 fields ASTRA's vector geometry would hold for this angle
        const double th = (double)angles[a];
        const double cs = cos(th), sn = sin(th);
        const float ray_x = (float)(-sn), ray_y = (float)cs;
        const float ux = (float)((double)det_width * cs), uy = (float)((double)det_width * sn);
        const float sx = (float)((double)det_width * -0.5 * (double)ndet * cs);
        const float sy = (float)((double)det_width * -0.5 * (double)ndet * sn);
        const int vertical = fabsf(ray_x) < fabsf(ray_y);
        FpAngle &f = fp[a];
        f.vertical = vertical;
        f.sx = sx; f.sy = sy; f.ux = ux; f.uy = uy;
        float ratio;
        if (vertical) {
            ratio = ray_x / ray_y;
            f.len = sqrtf(ray_y * ray_y + ray_x * ray_x) / fabsf(ray_y);
            // c = Dx + (Ey - Dy) * ratio - Ex,  Dx = sx + (d + .5) ux,  Dy = sy + (d + .5) uy
            f.c0_base = (double)sx + (Ey - (double)sy) * (double)ratio - Ex;
            f.c0_step = (double)ux - (double)uy * (double)ratio;
        } else {
            ratio = ray_y / ray_x;
            f.len = sqrtf(ray_y * ray_y + ray_x * ray_x) / fabsf(ray_x);
            // r = -(Dy + (Ex - Dx) * ratio - Ey)
            f.c0_base = -((double)sy + (Ex - (double)sx) * (double)ratio - Ey);
            f.c0_step = -((double)uy - (double)ux * (double)ratio);
        }
        f.delta = -ratio;
        f.ratio = ratio;
        pl.warp[a] = make_warp(f.delta, positions);
        // gather form: pixel (line l, position k) sits at detector coordinate d* = (k - c0_base - l*delta)/c0_step - 0.5
        BpAngle &b = bp[a];
        const double inv = 1.0 / f.c0_step;
        const double dk = inv, dl = -(double)f.delta * inv;
        b.d00 = -f.c0_base * inv - 0.5;
        b.dcol = vertical ? dk : dl;
        b.drow = vertical ? dl : dk;
        b.a0 = f.len;
        b.b = (float)((double)f.len * fabs(f.c0_step));
        b.a1 = (float)((double)f.len - (double)f.len * fabs(f.c0_step));
        b.pad_ = 0.0f;
        // float views of the warp for the kernels' searches
        const WarpAngle &w = pl.warp[a];
        WarpAngleF &wf = pl.warpf[a];
        for (int i = 0; i < 16; ++i) { wf.gcum[i] = i < WT_NB ? (float)w.gcum[i] : 3.0e38f; wf.islope[i] = i < WT_NB ? (float)w.islope[i] : 1.0f; }
        BpWarpF &bw = pl.bwarp[a];
        const double s0 = w.gcum[5] / 64.0;
        for (int m = 0; m < WT_NM; ++m) {
            bw.cum[m] = m ? (float)w.gcum[m + 4] : 0.0f;
            const double isl = m ? w.islope[m + 4] : 1.0 / s0, isl_prev = m > 1 ? w.islope[m + 3] : 1.0 / s0;
            bw.bend[m] = m ? (float)fabs(isl - isl_prev) : 0.0f;
        }
        bw.a0 = b.a0;
        bw.b = b.b;
        for (int m = 0; m < WT_NM; ++m) {
            bw.slope[m] = (float)(m ? w.slope[m + 4] : s0);
            bw.isl[m] = (float)(m ? w.islope[m + 4] : 1.0 / s0);
        }
        bw.inv_step = (float)(1.0 / f.c0_step);
        bw.delta = f.delta;
        BpWarpD &bd = pl.bwarpd[a];
        for (int m = 0; m < WT_NM; ++m) {
            bd.cum[m] = m ? w.gcum[m + 4] : 0.0;
            bd.slope[m] = m ? w.slope[m + 4] : s0;
            bd.isl[m] = m ? w.islope[m + 4] : 1.0 / s0;
        }
        bd.c0_base = f.c0_base; bd.inv_step = 1.0 / f.c0_step; bd.delta = (double)f.delta;
        bd.vertical = vertical; bd.pad_ = 0;
    }
    // angle tiles of the forward projector: consecutive angles of one orientation whose rays, over any
    // FP_DT-detector x FP_LB-line patch, stay within FP_WFIT positions of each other.  In exact arithmetic c_a - c_b is
    // affine in (d, l), so the spread across angles peaks at a corner of the (d, l) rectangle; the reference's running
    // sums bend the rays by up to a few positions (2.5 at 8192^2), piecewise linearly, so the spread is sampled on a
    // 5 x 5 grid of the warped positions.
    std::vector<FpTile> &tiles = pl.tiles;
    tiles.clear();
    {
        auto c_at = [&](int a, double d, double l) {
            const double c0 = fp[a].c0_base + (d + 0.5) * fp[a].c0_step;
            return warp_ginv(pl.warp[a], warp_g(pl.warp[a], c0) + l * (double)fp[a].delta);
        };
        // a single angle's own spread over a CTA's patch must fit (wide detector pixels near 45 degrees do not),
        // unless the staged window holds every in-volume position of a line wherever it starts
        for (int a = 0; a < nangles; ++a)
            if ((fp[a].vertical ? cols : rows) > FP_W - VPAD - 1 &&
                (FP_DT - 1) * fabs(fp[a].c0_step) + (FP_LB - 1) * fabs((double)fp[a].delta) > (double)FP_WFIT) {
                g_err = "geometry: detector pixels too wide for the forward projector's staging window (det_width <~ 1.6)";
                return WT_ERR_UNSUPPORTED;
            }
        int a = 0;
        while (a < nangles) {
            int cnt = 1;
            while (cnt < FP_AT && a + cnt < nangles && fp[a + cnt].vertical == fp[a].vertical) {
                const int n = cnt + 1;
                const int nl = fp[a].vertical ? rows : cols;
                double worst = 0.0, ms = 0.0, md = 0.0;
                for (int i = 0; i < n; ++i) { ms = fmax(ms, fabs(fp[a + i].c0_step)); md = fmax(md, fabs((double)fp[a + i].delta)); }
                for (int di = 0; di < 5; ++di)
                    for (int li = 0; li < 5; ++li) {
                        double lo = 1e300, hi = -1e300;
                        for (int i = 0; i < n; ++i) {
                            const double c = c_at(a + i, 0.25 * di * (ndet - 1), 0.25 * li * (nl - 1));
                            lo = fmin(lo, c); hi = fmax(hi, c);
                        }
                        worst = fmax(worst, hi - lo);
                    }
                if (worst + (FP_DT - 1) * ms + (FP_LB - 1) * md > (double)FP_WFIT) break;
                cnt = n;
            }
            tiles.push_back({a, cnt});
            a += cnt;
        }
    }
    // launch order: one phase per orientation (fp.cuh: fp_block); WT_FP_PHASES = 1 | 2 | 4 for experiments
    // (1: all tiles in one phase, 4: orientation x slope sign)
    std::vector<int> &order = pl.order;
    order.clear();
    int (&phase_start)[FP_MAX_PHASES + 1] = pl.phase_start;
    for (int i = 0; i <= FP_MAX_PHASES; ++i) phase_start[i] = 0;
    int &nphase = pl.nphase;
    nphase = 0;
    {
        const char *e = getenv("WT_FP_PHASES");
        int want = e ? atoi(e) : 2;
        if (want != 1 && want != 4) want = 2;
        for (int ph = 0; ph < 4; ++ph) {
            for (int t = 0; t < (int)tiles.size(); ++t) {
                const FpAngle &f = fp[tiles[t].first + tiles[t].count / 2];
                const int key = (f.vertical ? 0 : 2) + (f.delta < 0.0f ? 1 : 0);
                const int mine = want == 4 ? key : want == 2 ? (key & 2) : 0;
                if (mine == ph) order.push_back(t);
            }
            if ((int)order.size() > phase_start[nphase]) { ++nphase; phase_start[nphase] = (int)order.size(); }
        }
        for (int i = nphase + 1; i <= FP_MAX_PHASES; ++i) phase_start[i] = phase_start[nphase];
    }
    return WT_OK;
}

}  // namespace wt

using namespace wt;

extern "C" {

const char *wt_last_error(void) { return g_err.c_str(); }
int wt_version(void) { return 100; }

long long wt_debug_oob_count(void) {
    unsigned long long n = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(&n, g_oob_count, sizeof(n)) != cudaSuccess) return -1;
    return (long long)n;
}

int wt_geom_create(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, wt_geom **out) {
    return wt_geom_create_ex(rows, cols, ndet, det_width, angles, nangles, default_positions(), out);
}

int wt_geom_positions(const wt_geom *g) { return g ? g->positions : WT_ERR_INVALID; }

int wt_geom_create_ex(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, int positions,
                      wt_geom **out) {
    if (!out) return invalid("wt_geom_create: null pointer");
    if (positions != WT_POS_ASTRA && positions != WT_POS_EXACT) return invalid("wt_geom_create: bad positions mode");
    if (int rc = check_geometry(rows, cols, ndet, det_width, angles, nangles, "wt_geom_create")) return rc;
    GeomPlan pl;
    if (int rc = plan_geometry(rows, cols, ndet, det_width, angles, nangles, positions, pl)) return rc;
    const std::vector<FpAngle> &fp = pl.fp;
    const std::vector<BpAngle> &bp = pl.bp;
    const std::vector<FpTile> &tiles = pl.tiles;
    const std::vector<int> &order = pl.order;
    const int nphase = pl.nphase;
    const int (&phase_start)[FP_MAX_PHASES + 1] = pl.phase_start;
    wt_geom *g = new wt_geom();
    g->rows = rows; g->cols = cols; g->ndet = ndet; g->nangles = nangles; g->det_width = det_width;
    g->positions = positions;
    g->step_ev = nullptr; g->step_count = 0;
    // zero margins of the packed sinogram: every staged segment of every tile must stay inside a row (the reference's
    // running sums displace a pixel's detector coordinate by up to N * ulp(2N) / 2: 0.25 at 2048^2, 4 at 8192^2)
    const double halfdiag = 0.5 * sqrt((double)rows * rows + (double)cols * cols) / (double)det_width;
    const double nmax = (double)(rows > cols ? rows : cols);
    const int drift = positions == WT_POS_ASTRA ? (int)ceil(nmax * ldexp(2.0 * nmax, -24)) + 1 : 0;
    int padl = (int)ceil(halfdiag - (0.5 * ndet - 0.5)) + 3 + drift;
    if (padl < 4) padl = 4;
    padl = (padl + 3) & ~3;
    g->sino_padl = padl;
    g->sino_pitch = (padl + ndet + padl + BP_SEGW_MAX + 3) & ~3;
    g->d_fp = nullptr; g->d_bp = nullptr; g->d_warp = nullptr; g->d_warpf = nullptr; g->d_bwarp = nullptr; g->d_bwarpd = nullptr; g->d_bplans = nullptr; g->d_tiles = nullptr; g->d_order = nullptr;
    g->nphase = nphase;
    for (int i = 0; i <= FP_MAX_PHASES; ++i) g->phase_start[i] = phase_start[i];
    g->ntiles = (int)tiles.size();
    g->h_angles = (float *)malloc(sizeof(float) * nangles);
    memcpy(g->h_angles, angles, sizeof(float) * nangles);
    g->h_vertical = (unsigned char *)malloc(nangles);
    for (int a = 0; a < nangles; ++a) g->h_vertical[a] = (unsigned char)fp[a].vertical;
    cudaError_t e = cudaMalloc(&g->d_fp, sizeof(FpAngle) * nangles);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_bp, sizeof(BpAngle) * nangles);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_fp, fp.data(), sizeof(FpAngle) * nangles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_bp, bp.data(), sizeof(BpAngle) * nangles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_warp, sizeof(WarpAngle) * nangles);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_warp, pl.warp.data(), sizeof(WarpAngle) * nangles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_warpf, sizeof(WarpAngleF) * nangles);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_warpf, pl.warpf.data(), sizeof(WarpAngleF) * nangles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_bwarp, sizeof(BpWarpF) * nangles);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_bwarp, pl.bwarp.data(), sizeof(BpWarpF) * nangles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_bwarpd, sizeof(BpWarpD) * nangles);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_bwarpd, pl.bwarpd.data(), sizeof(BpWarpD) * nangles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_tiles, sizeof(FpTile) * tiles.size());
    if (e == cudaSuccess) e = cudaMemcpy(g->d_tiles, tiles.data(), sizeof(FpTile) * tiles.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_order, sizeof(int) * order.size());
    if (e == cudaSuccess) e = cudaMemcpy(g->d_order, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice);
    // small geometries: K2's plans for every (tile, angle), computed once by the function the kernel otherwise runs inline
    {
        const int th = bp_tile_height(rows, cols);
        const size_t ntx = (cols + BP_TW - 1) / BP_TW, nty = (rows + th - 1) / th;
        const size_t bytes = ntx * nty * (size_t)nangles * sizeof(BpPlanRec);
        static const int off = getenv("WT_NO_PLAN_TABLE") ? 1 : 0;      // A/B measurements
        if (e == cudaSuccess && !off && bytes <= WT_BP_PLAN_TABLE_MAX) {
            e = cudaMalloc(&g->d_bplans, bytes);
            if (e == cudaSuccess) {
                dim3 grid((unsigned)ntx, (unsigned)nty, (nangles + 127) / 128);
                if (th == 16) bp_plan_table_kernel<16><<<grid, 128>>>(g->d_bwarp, g->d_bwarpd, padl, rows, cols, nangles, (BpPlanRec *)g->d_bplans);
                else bp_plan_table_kernel<32><<<grid, 128>>>(g->d_bwarp, g->d_bwarpd, padl, rows, cols, nangles, (BpPlanRec *)g->d_bplans);
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
            }
        }
    }
    if (e != cudaSuccess) {
        wt_geom_destroy(g);
        return cuda_fail(e, "wt_geom_create");
    }
    *out = g;
    return WT_OK;
}

void wt_geom_destroy(wt_geom *g) {
    if (!g) return;
    if (g->d_fp) cudaFree(g->d_fp);
    if (g->d_bp) cudaFree(g->d_bp);
    if (g->d_warp) cudaFree(g->d_warp);
    if (g->d_warpf) cudaFree(g->d_warpf);
    if (g->d_bwarp) cudaFree(g->d_bwarp);
    if (g->d_bwarpd) cudaFree(g->d_bwarpd);
    if (g->d_bplans) cudaFree(g->d_bplans);
    if (g->d_tiles) cudaFree(g->d_tiles);
    if (g->d_order) cudaFree(g->d_order);
    if (g->step_ev) {
        for (int i = 0; i < 3 * WT_STEP_RING; ++i) cudaEventDestroy(g->step_ev[i]);
        free(g->step_ev);
    }
    free(g->h_angles);
    free(g->h_vertical);
    delete g;
}

int wt_geom_dims(const wt_geom *g, int *rows, int *cols, int *ndet, int *nangles) {
    if (!g) return invalid("wt_geom_dims: null geometry");
    if (rows) *rows = g->rows;
    if (cols) *cols = g->cols;
    if (ndet) *ndet = g->ndet;
    if (nangles) *nangles = g->nangles;
    return WT_OK;
}

int wt_geom_plan(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, int max_tiles,
                 int *tile_first, int *tile_count, int *tile_vertical, int *launch_order, int *n_tiles, int *n_phases,
                 int *phase_start) {
    if (!n_tiles) return invalid("wt_geom_plan: null pointer");
    if (int rc = check_geometry(rows, cols, ndet, det_width, angles, nangles, "wt_geom_plan")) return rc;
    GeomPlan pl;
    if (int rc = plan_geometry(rows, cols, ndet, det_width, angles, nangles, default_positions(), pl)) return rc;
    const int nt = (int)pl.tiles.size();
    *n_tiles = nt;
    if (n_phases) *n_phases = pl.nphase;
    if (phase_start) for (int i = 0; i <= FP_MAX_PHASES; ++i) phase_start[i] = pl.phase_start[i];
    for (int t = 0; t < nt && t < max_tiles; ++t) {
        if (tile_first) tile_first[t] = pl.tiles[t].first;
        if (tile_count) tile_count[t] = pl.tiles[t].count;
        if (tile_vertical) tile_vertical[t] = pl.fp[pl.tiles[t].first].vertical;
        if (launch_order) launch_order[t] = pl.order[t];
    }
    return WT_OK;
}

int wt_split_plan(int rows, int cols, int ndet, int nangles, int n_tiles, int *fp_split, int *bp_split, int *bp_tile_rows) {
    if (rows <= 0 || cols <= 0 || ndet <= 0 || nangles <= 0 || n_tiles <= 0) return invalid("wt_split_plan: non-positive size");
    if (fp_split) *fp_split = fp_cluster_split(((ndet + FP_DT - 1) / FP_DT) * n_tiles, rows > cols ? rows : cols);
    if (bp_split) *bp_split = bp_cluster_split(rows, cols, nangles);
    if (bp_tile_rows) *bp_tile_rows = bp_tile_height(rows, cols);
    return WT_OK;
}

#if WT_TRACE
// development builds only (not in the header): copy the per-CTA time stamps of the last K1 (kernel 0) / K2 (kernel 1) launch
int wt_debug_trace(int kernel, unsigned long long *out, int nctas) {
    if (kernel < 0 || kernel > 1 || nctas > WT_TRACE_CTAS) return WT_ERR_INVALID;
    cudaDeviceSynchronize();
    WT_CUDA(cudaMemcpyFromSymbol(out, g_wt_trace, sizeof(unsigned long long) * WT_TRACE_POINTS * nctas,
                                 sizeof(unsigned long long) * WT_TRACE_POINTS * WT_TRACE_CTAS * kernel));
    return WT_OK;
}
#endif

size_t wt_workspace_bytes(const wt_geom *g, int nimages) {
    if (!g || nimages <= 0) return 0;
    nimages = packed_images(g, nimages);   // one or two images are packed together with their point reflections
    size_t fl = packed_vol_floats(g, nimages) + packed_sino_floats(g, nimages);
    fl += align_up((size_t)3 * nimages * fp_blocks(g), 64);                  // per-CTA partials (float)
    size_t bytes = fl * sizeof(float) + align_up((size_t)nimages * UPD_BLOCKS_PER_SLICE * sizeof(double), 256);
    return bytes + 4096;
}

int wt_fp(const wt_geom *g, const float *vol, float *sino, int nimages, int angle_begin, int angle_end, void *ws,
          size_t ws_bytes, void *stream) {
    if (!g || !vol || !sino || !ws) return invalid("wt_fp: null pointer");
    if (nimages <= 0) return invalid("wt_fp: nimages must be positive");
    if (angle_begin < 0 || angle_end > g->nangles || angle_begin > angle_end) return invalid("wt_fp: bad angle range");
    if (ws_bytes < wt_workspace_bytes(g, nimages)) { g_err = "wt_fp: workspace too small"; return WT_ERR_WORKSPACE; }
    const int npack = packed_images(g, nimages);
    Carver cv(ws, ws_bytes);
    float *vn = cv.take<float>(vn_floats(g, npack));
    float *vt = cv.take<float>(vt_floats(g, npack));
    if (!cv.ok()) { g_err = "wt_fp: workspace too small"; return WT_ERR_WORKSPACE; }
    HFpStore epi;
    epi.sino = sino;
    const int rc = run_fp(g, vol, nimages, angle_begin, angle_end, vn, vt, epi, (cudaStream_t)stream);
    return rc < 0 ? rc : WT_OK;
}

int wt_mar_fp(const wt_geom *g, const float *vol, const float *b, const unsigned char *mask, const float *inv_ngood,
              float bt, float b_hb, float *u, float *proj, double *stats, int nimages, void *ws, size_t ws_bytes,
              void *stream) {
    if (!g || !vol || !b || !mask || !inv_ngood || !ws) return invalid("wt_mar_fp: null pointer");
    if (nimages <= 0) return invalid("wt_mar_fp: nimages must be positive");
    if (ws_bytes < wt_workspace_bytes(g, nimages)) { g_err = "wt_mar_fp: workspace too small"; return WT_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(ws, ws_bytes);
    const int npack = packed_images(g, nimages);
    float *vn = cv.take<float>(vn_floats(g, npack));
    float *vt = cv.take<float>(vt_floats(g, npack));
    cv.take<float>((size_t)npack * g->nangles * g->sino_pitch);
    float *partial = cv.take<float>((size_t)3 * npack * fp_blocks(g));
    if (!cv.ok()) { g_err = "wt_mar_fp: workspace too small"; return WT_ERR_WORKSPACE; }
    HFpMar e;
    e.b = b; e.mask = mask; e.inv_ngood = inv_ngood; e.bt = bt; e.b_hb = b_hb; e.u = u; e.proj = proj;
    e.partial = stats ? partial : nullptr; e.nimg = nimages;
    const int nb = run_fp(g, vol, nimages, 0, g->nangles, vn, vt, e, st);   // partial-sum slots per image
    if (nb < 0) return nb;
    if (stats) {
        // stats is (nimages, 3): reduce each of the three planes with stride 3
        for (int q = 0; q < 3; ++q) {
            reduce_partials_strided_kernel<<<nimages, 32, 0, st>>>(partial + (size_t)q * nimages * nb, nb, stats + q, 3);
            WT_CUDA(cudaGetLastError());
        }
    }
    return WT_OK;
}

int wt_bp(const wt_geom *g, const float *sino, float *vol, int nimages, int angle_begin, int angle_end, float scale,
          void *ws, size_t ws_bytes, void *stream) {
    return wt_bp_rows(g, sino, vol, nimages, angle_begin, angle_end, 0, -1, 0, scale, ws, ws_bytes, stream);
}

int wt_bp_row_info(const wt_geom *g, int nimages, int *rows_processed, int *row_align, int *mirrored) {
    if (!g || nimages <= 0) return invalid("wt_bp_row_info: bad arguments");
    const bool m = use_mirror(g, nimages);
    if (rows_processed) *rows_processed = m ? (g->rows + 1) / 2 : g->rows;
    if (row_align) *row_align = bp_tile_height(g->rows, g->cols);
    if (mirrored) *mirrored = m ? 1 : 0;
    return WT_OK;
}

int wt_bp_rows(const wt_geom *g, const float *sino, float *vol, int nimages, int angle_begin, int angle_end, int row_begin,
               int row_end, int reuse_packed, float scale, void *ws, size_t ws_bytes, void *stream) {
    if (!g || !vol || !sino || !ws) return invalid("wt_bp: null pointer");
    if (nimages <= 0) return invalid("wt_bp: nimages must be positive");
    if (angle_begin < 0 || angle_end > g->nangles || angle_begin > angle_end) return invalid("wt_bp: bad angle range");
    if (ws_bytes < wt_workspace_bytes(g, nimages)) { g_err = "wt_bp: workspace too small"; return WT_ERR_WORKSPACE; }
    const int npack = packed_images(g, nimages);
    Carver cv(ws, ws_bytes);
    float *sp = cv.take<float>((size_t)npack * g->nangles * g->sino_pitch);
    if (!cv.ok()) { g_err = "wt_bp: workspace too small"; return WT_ERR_WORKSPACE; }
    HBpStore epi;
    epi.vol = vol;
    epi.scale = scale;
    return run_bp(g, sino, nimages, angle_begin, angle_end, sp, epi, (cudaStream_t)stream, row_begin, row_end, reuse_packed != 0);
}

size_t wt_sirt_workspace_bytes(const wt_geom *g, int nimages) {
    if (!g || nimages <= 0) return 0;
    const size_t plane = (size_t)g->nangles * g->ndet, img = (size_t)g->rows * g->cols;
    return wt_workspace_bytes(g, nimages) + (align_up(plane, 64) + align_up(img, 64) + align_up((size_t)nimages * plane, 64)) * sizeof(float) + 1024;
}

int wt_sirt(const wt_geom *g, const float *sino, float *vol, int nimages, int n_iters, void *ws, size_t ws_bytes,
            void *stream) {
    if (!g || !vol || !sino || !ws) return invalid("wt_sirt: null pointer");
    if (nimages <= 0 || n_iters < 0) return invalid("wt_sirt: bad counts");
    if (ws_bytes < wt_sirt_workspace_bytes(g, nimages)) { g_err = "wt_sirt: workspace too small"; return WT_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)g->nangles * g->ndet, img = (size_t)g->rows * g->cols;
    Carver cv(ws, ws_bytes);
    const int npack = packed_images(g, nimages);   // the one-image W 1 / W^T 1 passes never pack more than this
    float *vn = cv.take<float>(vn_floats(g, npack));
    float *vt = cv.take<float>(vt_floats(g, npack));
    float *sp = cv.take<float>((size_t)npack * g->nangles * g->sino_pitch);
    float *raysum = cv.take<float>(plane);
    float *pixsum = cv.take<float>(img);
    float *resid = cv.take<float>((size_t)nimages * plane);
    if (!cv.ok()) { g_err = "wt_sirt: workspace too small"; return WT_ERR_WORKSPACE; }
    // Rsum = W 1, Csum = W^T 1 (the residual buffer doubles as the all-ones input)
    fill_kernel<<<296, 256, 0, st>>>(vol, img, 1.0f);
    HFpStore fs; fs.sino = raysum;
    int rc = run_fp(g, vol, 1, 0, g->nangles, vn, vt, fs, st);
    if (rc < 0) return rc;
    fill_kernel<<<296, 256, 0, st>>>(resid, plane, 1.0f);
    HBpStore bs; bs.vol = pixsum; bs.scale = 1.0f;
    rc = run_bp(g, resid, 1, 0, g->nangles, sp, bs, st);
    if (rc) return rc;
    WT_CUDA(cudaMemsetAsync(vol, 0, sizeof(float) * nimages * img, st));
    // Two launches per iteration: K1's epilogue writes the weighted residual straight into K2's packed sinogram layout,
    // K2's epilogue updates the volume and refreshes K1's packed copies (the zero padding of the packed buffers is
    // written once, here; x0 = 0 is all zeros).  Mirror-paired calls (WT_POS_EXACT, one or two large images) keep the
    // pack kernels: their packed copies also hold the point reflections.
    const bool fused = !use_mirror(g, nimages);
    if (fused) {
        WT_CUDA(cudaMemsetAsync(vn, 0, sizeof(float) * vn_floats(g, npack), st));
        WT_CUDA(cudaMemsetAsync(vt, 0, sizeof(float) * vt_floats(g, npack), st));
        WT_CUDA(cudaMemsetAsync(sp, 0, sizeof(float) * (size_t)npack * g->nangles * g->sino_pitch, st));
    }
    for (int it = 0; it < n_iters; ++it) {
        HFpSirt fe; fe.meas = sino; fe.raysum = raysum; fe.out = fused ? nullptr : resid;
        fe.sp = fused ? sp : nullptr; fe.padl = g->sino_padl; fe.pitch = g->sino_pitch;
        rc = run_fp(g, vol, nimages, 0, g->nangles, vn, vt, fe, st, fused);
        if (rc < 0) return rc;
        HBpSirt be; be.vol = vol; be.pixsum = pixsum; be.vn = fused ? vn : nullptr; be.vt = fused ? vt : nullptr;
        rc = run_bp(g, resid, nimages, 0, g->nangles, sp, be, st, 0, -1, fused);
        if (rc) return rc;
    }
    return WT_OK;
}

int wt_spectrum_create(int n_elements, int n_bins, const double *source, const double *widths, const double *absorptions,
                       double pixel_size, wt_spectrum **out) {
    if (!out || !source || !widths || !absorptions) return invalid("wt_spectrum_create: null pointer");
    if (n_elements <= 0 || n_bins <= 0) return invalid("wt_spectrum_create: sizes must be positive");
    std::vector<float> tab((size_t)(n_elements + 1) * n_bins);
    for (int e = 0; e < n_bins; ++e) tab[e] = (float)(source[e] * widths[e]);
    for (int k = 0; k < n_elements; ++k)
        for (int e = 0; e < n_bins; ++e) tab[(size_t)(k + 1) * n_bins + e] = (float)(pixel_size * absorptions[(size_t)k * n_bins + e] * 1.4426950408889634);
    wt_spectrum *s = new wt_spectrum();
    s->K = n_elements; s->E = n_bins; s->d_tab = nullptr;
    s->cslot = -1; s->device = 0;
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_tab, tab.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && tab.size() <= (size_t)WT_CTAB) {
        std::lock_guard<std::mutex> lock(g_slot_mutex);
        unsigned &used = g_slots_used[s->device & 63];
        for (int i = 0; i < WT_CSLOTS && s->cslot < 0; ++i)
            if (!(used & (1u << i))) { used |= 1u << i; s->cslot = i; }
        // a kernel of a destroyed spectrum may still be reading the slot on some stream: creation is set-up time, wait
        if (s->cslot >= 0) e = cudaDeviceSynchronize();
        if (s->cslot >= 0 && e == cudaSuccess)
            e = cudaMemcpyToSymbol(c_spectrum_tab, tab.data(), tab.size() * sizeof(float), (size_t)s->cslot * WT_CTAB * sizeof(float));
    }
    if (e != cudaSuccess) { wt_spectrum_destroy(s); return cuda_fail(e, "wt_spectrum_create"); }
    *out = s;
    return WT_OK;
}

void wt_spectrum_destroy(wt_spectrum *s) {
    if (!s) return;
    if (s->d_tab) cudaFree(s->d_tab);
    if (s->cslot >= 0) {
        std::lock_guard<std::mutex> lock(g_slot_mutex);
        g_slots_used[s->device & 63] &= ~(1u << s->cslot);
    }
    delete s;
}

int wt_white_fp(const wt_geom *g, const wt_spectrum *s, const float *conc, const float *meas, float *intensity,
                float *qmu, double *loss, int nslices, int angle_begin, int angle_end, void *ws, size_t ws_bytes,
                void *stream) {
    if (!g || !s || !conc || !ws) return invalid("wt_white_fp: null pointer");
    if (nslices <= 0) return invalid("wt_white_fp: nslices must be positive");
    if (angle_begin < 0 || angle_end > g->nangles || angle_begin > angle_end) return invalid("wt_white_fp: bad angle range");
    const int K = s->K, n = nslices * K;
    {
        int dev = -1;
        WT_CUDA(cudaGetDevice(&dev));
        if (dev != s->device) return invalid("wt_white_fp: the spectrum was created on another device than the current one");
    }
    if (ws_bytes < wt_workspace_bytes(g, n)) { g_err = "wt_white_fp: workspace too small"; return WT_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)g->nangles * g->ndet;
    Carver cv(ws, ws_bytes);
    const int npack = packed_images(g, n);
    float *vn = cv.take<float>(vn_floats(g, npack));
    float *vt = cv.take<float>(vt_floats(g, npack));
    float *sp = cv.take<float>((size_t)npack * g->nangles * g->sino_pitch);  // generic-K path: holds P (n, A, D)
    float *partial = cv.take<float>((size_t)npack * fp_blocks(g));
    if (!cv.ok()) { g_err = "wt_white_fp: workspace too small"; return WT_ERR_WORKSPACE; }
    int nb = 0;
    if (K == 1 || K == 2) {
        nb = fp_blocks(g);
        int rc;
        if (K == 1) {
            HFpWhite<1> e; e.tab = s->d_tab; e.E = s->E; e.cslot = s->cslot; e.meas = meas; e.inten = intensity; e.qmu = qmu;
            e.partial = loss ? partial : nullptr; e.sp = nullptr; e.padl = 0; e.pitch = 0;
            rc = run_fp(g, conc, n, angle_begin, angle_end, vn, vt, e, st);
        } else {
            HFpWhite<2> e; e.tab = s->d_tab; e.E = s->E; e.cslot = s->cslot; e.meas = meas; e.inten = intensity; e.qmu = qmu;
            e.partial = loss ? partial : nullptr; e.sp = nullptr; e.padl = 0; e.pitch = 0;
            rc = run_fp(g, conc, n, angle_begin, angle_end, vn, vt, e, st);
        }
        if (rc < 0) return rc;
        nb = rc;   // partial-sum slots per slice (CTAs x passes)
    } else {
        HFpStore fs; fs.sino = sp;
        int rc = run_fp(g, conc, n, angle_begin, angle_end, vn, vt, fs, st);
        if (rc < 0) return rc;
        nb = (int)((plane + 255) / 256);
        white_epilogue_kernel<<<dim3(nb, nslices), 256, 0, st>>>(sp, s->d_tab, K, s->E, meas, intensity, qmu,
                                                                  loss ? partial : nullptr, plane,
                                                                  (size_t)angle_begin * g->ndet, (size_t)angle_end * g->ndet);
        WT_CUDA(cudaGetLastError());
    }
    if (loss) {
        reduce_partials_kernel<<<nslices, 32, 0, st>>>(partial, nb, 0.5, loss);
        WT_CUDA(cudaGetLastError());
    }
    return WT_OK;
}

size_t wt_step_workspace_bytes(const wt_geom *g, int nslices, int n_elements) {
    if (!g || nslices <= 0 || n_elements <= 0) return 0;
    return wt_workspace_bytes(g, nslices * n_elements) + align_up((size_t)nslices * bp_slots(g) * sizeof(double), 256) + 1024;
}

}  // extern "C"

template <int K>
static int white_barrier_step(const wt_geom *g, const wt_spectrum *s, float *x, const float *meas, const wt_update_params *prm,
                              const unsigned char *active, double *loss, double *penalty, int nslices, int flags, void *ws,
                              size_t ws_bytes, cudaStream_t st) {
    const int n = nslices * K;
    Carver cv(ws, ws_bytes);
    const int npack = packed_images(g, n);
    float *vn = cv.take<float>(vn_floats(g, npack));
    float *vt = cv.take<float>(vt_floats(g, npack));
    float *sp = cv.take<float>((size_t)npack * g->nangles * g->sino_pitch);
    float *lpart = cv.take<float>((size_t)3 * npack * fp_blocks(g));
    cv.take<double>((size_t)n * UPD_BLOCKS_PER_SLICE);
    const int nslots = bp_slots(g);
    double *ppart = cv.take<double>((size_t)nslices * nslots);
    if (!cv.ok()) { g_err = "wt_white_barrier_step: workspace too small (wt_step_workspace_bytes)"; return WT_ERR_WORKSPACE; }
    const Chunk c = chunk_of(n);
    cudaEvent_t *ev = nullptr;
    if ((flags & WT_STEP_TIMED) && !(flags & WT_STEP_INIT)) {
        wt_geom *gm = const_cast<wt_geom *>(g);
        if (!gm->step_ev) {
            gm->step_ev = (cudaEvent_t *)malloc(sizeof(cudaEvent_t) * 3 * WT_STEP_RING);
            for (int i = 0; i < 3 * WT_STEP_RING; ++i) WT_CUDA(cudaEventCreate(&gm->step_ev[i]));
        }
        ev = gm->step_ev + 3 * (gm->step_count % WT_STEP_RING);
        gm->step_count++;
        WT_CUDA(cudaEventRecord(ev[0], st));
    }
    if (flags & WT_STEP_INIT) {
        // packed copies of x (with their zero padding) and an all-zero packed sinogram buffer (K1 only writes detectors
        // [0, ndet) of every row); no update: the forward model below then produces q*mu and f(x) of the start point
        switch (c.v) {
        case 4: launch_pack_vol<4>(x, vn, vt, g->rows, g->cols, c.groups, n, 3, st); break;
        case 2: launch_pack_vol<2>(x, vn, vt, g->rows, g->cols, 1, n, 3, st); break;
        default: launch_pack_vol<1>(x, vn, vt, g->rows, g->cols, 1, n, 3, st); break;
        }
        WT_CUDA(cudaMemsetAsync(sp, 0, sizeof(float) * (size_t)npack * g->nangles * g->sino_pitch, st));
    } else {
        // K2 over the packed q*mu of the previous forward model; its epilogue takes the step and refreshes VN / VT
        BpEpiBarrierUpdate<K> be;
        be.x = x; be.prm = *prm; be.active = active; be.partial = penalty ? ppart : nullptr; be.vn = vn; be.vt = vt; be.gscale = -1.0f;
        const int rc = run_bp(g, nullptr, n, 0, g->nangles, sp, be, st, 0, -1, true);
        if (rc) return rc;
        if (ev) WT_CUDA(cudaEventRecord(ev[1], st));
    }
    HFpWhite<K> e;
    e.tab = s->d_tab; e.E = s->E; e.cslot = s->cslot; e.meas = meas; e.inten = nullptr; e.qmu = nullptr;
    e.partial = loss ? lpart : nullptr; e.sp = sp; e.padl = g->sino_padl; e.pitch = g->sino_pitch;
    const int nb = run_fp(g, x, n, 0, g->nangles, vn, vt, e, st, true);
    if (nb < 0) return nb;
    if (ev) WT_CUDA(cudaEventRecord(ev[2], st));
    if (loss || (penalty && !(flags & WT_STEP_INIT))) {
        reduce_step_kernel<<<dim3(nslices, 2), 32, 0, st>>>(lpart, nb, loss, (flags & WT_STEP_INIT) ? nullptr : ppart, nslots,
                                                             (flags & WT_STEP_INIT) ? nullptr : penalty);
        WT_CUDA(cudaGetLastError());
    }
    return WT_OK;
}

extern "C" {

int wt_step_times(const wt_geom *g, int max_steps, float *bp_ms, float *fp_ms) {
    if (!g || !bp_ms || !fp_ms || max_steps < 0) return WT_ERR_INVALID;
    if (!g->step_ev) return 0;
    const long long have = g->step_count < WT_STEP_RING ? g->step_count : WT_STEP_RING;
    const int n = (int)(have < max_steps ? have : max_steps);
    for (int i = 0; i < n; ++i) {        // the n most recent timed steps, oldest first
        const cudaEvent_t *ev = g->step_ev + 3 * ((g->step_count - n + i) % WT_STEP_RING);
        if (cudaEventSynchronize(ev[2]) != cudaSuccess) return WT_ERR_CUDA;
        if (cudaEventElapsedTime(&bp_ms[i], ev[0], ev[1]) != cudaSuccess) return WT_ERR_CUDA;
        if (cudaEventElapsedTime(&fp_ms[i], ev[1], ev[2]) != cudaSuccess) return WT_ERR_CUDA;
    }
    return n;
}

int wt_white_barrier_step(const wt_geom *g, const wt_spectrum *s, float *x, const float *meas, const wt_update_params *prm,
                          const unsigned char *active, double *loss, double *penalty, int nslices, int flags, void *ws,
                          size_t ws_bytes, void *stream) {
    if (!g || !s || !x || !ws) return invalid("wt_white_barrier_step: null pointer");
    if (!(flags & WT_STEP_INIT) && !prm) return invalid("wt_white_barrier_step: update parameters required");
    if (nslices <= 0) return invalid("wt_white_barrier_step: nslices must be positive");
    const int K = s->K;
    if (K != 1 && K != 2) { g_err = "wt_white_barrier_step: fused iteration needs K <= 2 (use wt_white_fp / wt_bp / wt_barrier_update)"; return WT_ERR_UNSUPPORTED; }
    if (use_mirror(g, nslices * K)) { g_err = "wt_white_barrier_step: not available for mirror-paired calls (WT_POS_EXACT, <= 2 large images)"; return WT_ERR_UNSUPPORTED; }
    if (prm && prm->reg_w != 0.0f && K != 2) return invalid("wt_white_barrier_step: the |c0 c1| regulariser needs K == 2");
    if (prm && (prm->mode < 0 || prm->mode > 2)) return invalid("wt_white_barrier_step: bad mode");
    {
        int dev = -1;
        WT_CUDA(cudaGetDevice(&dev));
        if (dev != s->device) return invalid("wt_white_barrier_step: the spectrum was created on another device than the current one");
    }
    if (ws_bytes < wt_step_workspace_bytes(g, nslices, K)) { g_err = "wt_white_barrier_step: workspace too small"; return WT_ERR_WORKSPACE; }
    if (K == 1) return white_barrier_step<1>(g, s, x, meas, prm, active, loss, penalty, nslices, flags, ws, ws_bytes, (cudaStream_t)stream);
    return white_barrier_step<2>(g, s, x, meas, prm, active, loss, penalty, nslices, flags, ws, ws_bytes, (cudaStream_t)stream);
}

int wt_barrier_update(float *x, const float *grad, int nslices, int n_elements, size_t npix, const wt_update_params *prm,
                      const unsigned char *active, double *penalty, void *ws, size_t ws_bytes, void *stream) {
    if (!x || !prm) return invalid("wt_barrier_update: null pointer");
    if (!grad && prm->mode != WT_UPD_CLIP_ONLY) return invalid("wt_barrier_update: grad is required unless mode is WT_UPD_CLIP_ONLY");
    if (nslices <= 0 || n_elements <= 0 || n_elements > 4 || npix == 0) return invalid("wt_barrier_update: bad sizes (1 <= K <= 4)");
    if (prm->reg_w != 0.0f && n_elements != 2) return invalid("wt_barrier_update: the |c0 c1| regulariser needs K == 2");
    if (prm->mode < 0 || prm->mode > 2) return invalid("wt_barrier_update: bad mode");
    double *partial = nullptr;
    if (penalty) {
        if (!ws || ws_bytes < (size_t)nslices * UPD_BLOCKS_PER_SLICE * sizeof(double)) { g_err = "wt_barrier_update: workspace too small"; return WT_ERR_WORKSPACE; }
        partial = (double *)ws;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (npix % 4 == 0) && (((uintptr_t)x | (uintptr_t)grad) % 16 == 0);
    const size_t work = vec ? npix / 4 : npix;
    int bx = (int)((work + 255) / 256);
    if (bx > UPD_BLOCKS_PER_SLICE) bx = UPD_BLOCKS_PER_SLICE;
    dim3 grid(bx, nslices);
    const bool fast = vec && n_elements == 2 && prm->reg_w != 0.0f && prm->t != 0.0f && !prm->lower_only &&
                      prm->triv_w == 0.0f && (prm->elem_mask & 3u) == 3u && prm->mode == WT_UPD_STEP_CLIP;
#define WT_UPD_LAUNCH(KK)                                                                                     \
    if (vec) barrier_update_kernel<KK, 4, false><<<grid, 256, 0, st>>>(x, grad, npix, *prm, active, partial);  \
    else barrier_update_kernel<KK, 1, false><<<grid, 256, 0, st>>>(x, grad, npix, *prm, active, partial)
    switch (n_elements) {
    case 1: WT_UPD_LAUNCH(1); break;
    case 2:
        if (fast) barrier_update_kernel<2, 4, true><<<grid, 256, 0, st>>>(x, grad, npix, *prm, active, partial);
        else { WT_UPD_LAUNCH(2); }
        break;
    case 3: WT_UPD_LAUNCH(3); break;
    default: WT_UPD_LAUNCH(4); break;
    }
#undef WT_UPD_LAUNCH
    WT_CUDA(cudaGetLastError());
    if (penalty) {
        reduce_partials_f64_kernel<<<nslices, 32, 0, st>>>(partial, bx, 1.0, penalty);
        WT_CUDA(cudaGetLastError());
    }
    return WT_OK;
}

}  // extern "C"

// ---- ||c0 c1||^2 per slice (the loss term of src/white_rec.py:236-237) ------------------------------------------------
// grid (WT_PAIR_NORM_BLOCKS, nslices): block b reduces a contiguous slab of the slice in float64 (the product itself is a
// float32 product, like the reference's array arithmetic on float32 data), fixed order within the block.
__global__ void __launch_bounds__(256) pair_norm2_kernel(const float *__restrict__ c, size_t npix, double *__restrict__ out) {
    const float *c0 = c + (size_t)blockIdx.y * 2 * npix, *c1 = c0 + npix;
    const size_t per = (npix + gridDim.x - 1) / gridDim.x;
    const size_t i0 = (size_t)blockIdx.x * per, i1 = i0 + per < npix ? i0 + per : npix;
    double acc = 0.0;
    for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const float pr = __fmul_rn(__ldg(c0 + i), __ldg(c1 + i));
        acc += (double)pr * (double)pr;
    }
    acc = wt::warp_sum(acc);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

int wt_pair_norm2(const float *conc, int nslices, size_t npix, double *out, void *stream) {
    if (!conc || !out) return invalid("wt_pair_norm2: null pointer");
    if (nslices <= 0 || npix == 0) return invalid("wt_pair_norm2: bad sizes");
    pair_norm2_kernel<<<dim3(WT_PAIR_NORM_BLOCKS, nslices), 256, 0, (cudaStream_t)stream>>>(conc, npix, out);
    WT_CUDA(cudaGetLastError());
    return WT_OK;
}

