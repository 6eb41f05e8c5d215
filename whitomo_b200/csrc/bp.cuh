// K2: pixel-driven backprojector, the exact transpose of K1 in gather form:
//   v[row, col] = sum_a  w0 * q[a, d0] + w1 * q[a, d0 + 1],   d0 = floor(d*),  f = d* - d0,
//   w0 = max(0, A0 - B f),  w1 = max(0, A1 + B f)           (triangle of half-width m, height 1/m).
//
// A CTA owns a BP_TW x TH pixel tile (TH = 32, or 16 for small volumes) and keeps TH/4 pixels x V images per thread in
// registers.  For every block of BP_AB angles a dedicated producer warp computes, per angle, where the tile's shadow
// falls on the detector and asks the TMA unit (cp.async.bulk, one 1-D bulk copy per angle) to stage that 80- (72-)
// detector sinogram segment in shared memory.  A full/empty mbarrier pair per stage, BP_STAGES deep,
// lets the copies of later angle blocks overlap the gather of the current one and lets the 8 consumer
// warps run without ever synchronising with each other.
// Lanes run along image columns: their detector positions differ by |cos| <= 1, so the LDS of a warp
// touches at most 32 consecutive words -- conflict-free.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace wt {

// ---- packing: plain (n, A, D) -> SP (G, A, pitch, V) with `padl` zero detectors in front ------------
// MIRROR: the group holds V0 = V/2 images; components [V0, V) are the detector-reversed rows q[a][D-1-d]
// (pixel (R-1-row, C-1-col) sees detector coordinate (D-1) - d*, see the kernel's epilogue)
template <int V, bool MIRROR = false>
__global__ void __launch_bounds__(256) pack_sino_kernel(const float *__restrict__ sino, float *__restrict__ sp, int A,
                                                        int D, int padl, int pitch, int nimg, int a_begin) {
    static_assert(!MIRROR || V % 2 == 0, "mirror pairing needs an even interleave width");
    typedef typename VecOf<V>::T VecT;
    constexpr int V0 = MIRROR ? V / 2 : V;
    const int g = blockIdx.z, a = a_begin + blockIdx.y;      // only the angles the backprojection will read
    const size_t plane = (size_t)A * D;
    for (int dp = blockIdx.x * blockDim.x + threadIdx.x; dp < pitch; dp += gridDim.x * blockDim.x) {
        const int d = dp - padl;
        float v[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const bool rev = MIRROR && i >= V0;
            const size_t im = (size_t)g * V0 + (rev ? i - V0 : i);       // >= nimg: padding component of the last group
            v[i] = (d >= 0 && d < D && im < (size_t)nimg) ? __ldg(sino + im * plane + (size_t)a * D + (rev ? D - 1 - d : d)) : 0.0f;
        }
        reinterpret_cast<VecT *>(sp)[((size_t)g * A + a) * pitch + dp] = pack<V>(v);
    }
}

template <int V, bool MIRROR = false>
inline void launch_pack_sino(const float *sino, float *sp, const wt_geom *g, int groups, int nimg, int a_begin, int a_end,
                             cudaStream_t st) {
    if (a_end <= a_begin) return;
    dim3 grid((g->sino_pitch + 255) / 256, a_end - a_begin, groups);
    pack_sino_kernel<V, MIRROR><<<grid, 256, 0, st>>>(sino, sp, g->nangles, g->ndet, g->sino_padl, g->sino_pitch, nimg, a_begin);
}

struct BpParams {
    const float *sp;  // packed sinogram (G, A, pitch, V)
    const BpWarpF *bwarp;    // per angle: K2's float view of the position warp + the tap-weight constants a0, b
    const BpWarpD *bwarpd;   // per angle: the same in double + c0_base, 1/c0_step, delta (identity warp under WT_POS_EXACT)
    int R, C, A;
    int nimg;         // images of this launch; the last group is padded with zero components up to V
    int Rproc;        // rows that get a thread: R, or ceil(R/2) under mirror pairing
    int row_begin, row_end;   // of those, the rows this launch computes (row_begin is a multiple of the tile height)
    int padl, pitch;
    int a_begin, a_end;
    const struct BpPlanRec *plans;   // precomputed plans of a small geometry (bp_plan_table_kernel), or null
    int csplit;       // CTAs of a cluster that share a tile, each taking a range of the angle blocks (1: no cluster)
};

// plain store: vol (n, R, C)
struct BpEpiStore {
    static constexpr bool kHasSums = false;
    static constexpr int kMult = 1;      // images per problem: a group of V interleaved images must hold whole problems
    float *vol;
    float scale;
    template <int V>
    __device__ __forceinline__ void operator()(const BpParams &p, int g, int row, int col,
                                               const float (&val)[V]) const {
        const size_t img = (size_t)p.R * p.C;
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (g * V + i < p.nimg) vol[((size_t)g * V + i) * img + (size_t)row * p.C + col] = scale * val[i];
    }
};

// Of a consumer thread's 2 columns x BP_RPT/2 vertically adjacent pixel pairs, how many share their taps (bp_kernel:
// shared_pair).  A shared pair trades one LDS.128 (4 shared-memory wavefronts) for a few scalar instructions.  Without
// sharing the float4 kernel is wavefront-bound (ncu: l1tex 97.6 %, issue 63 %, 27.5 ms at 2048^2 x 1800 x 16 images);
// with all four pairs shared, the lower pixel located from the upper one, the FMAs packed (FFMA2) and three weights
// on three taps it takes 22.0 ms (with selects on the taps instead: 23.2 ms at l1tex 89 %, issue 71 %; 2 / 3 / 4
// shared pairs: 24.8 / 23.5 / 23.2 ms).  The float2 and float kernels (one or two images) are issue-bound to begin
// with -- sharing costs them 7-12 % -- and keep separate taps.
#ifndef WT_BP_NSHARE
#define WT_BP_NSHARE 4
#endif
#ifndef WT_BP_PAIRPIX
#define WT_BP_PAIRPIX 1      // float / float2 kernels: two pixels per pass through the packed fp32 pipe (bp_kernel: pixel_pair)
#endif
template <int V> struct BpShare { static constexpr int value = V >= 4 ? WT_BP_NSHARE : 0; };
// tile row of a consumer thread's r-th pixel row: vertically adjacent pairs 2*warp + {0, 1}, then 16 rows further down
__host__ __device__ constexpr int bp_row_of(int warp, int r) { return 2 * warp + 16 * (r >> 1) + (r & 1); }

template <int V, int TH>
struct BpSmem {
    typedef typename VecOf<V>::T VecT;
    VecT seg[BP_STAGES][BP_AB][BpTile<TH>::SEGW];
    float4 cst[BP_STAGES][BP_AB];  // affine piece A of d*: dcol, drow, base - 1/2;  Ah = A0 - B/2
    float4 cst2[BP_STAGES][BP_AB]; // affine piece B of d*: dcol, drow, base - 1/2;  Ah of piece B
    float2 cst3[BP_STAGES][BP_AB]; // B of piece A, flags (BP_F_*)
    float cst4[BP_STAGES][BP_AB];  // B of piece B
    uint64_t full[BP_STAGES];   // TMA bytes of a stage have landed
    uint64_t empty[BP_STAGES];  // all consumer warps are done reading a stage
};

// How the consumers treat an angle of a stage (uniform over the CTA):
//   0            one affine piece, rays at least a position apart: two taps per pixel, pairs share taps (the fast path)
//   BP_F_KINK    two pieces whose slopes differ by < 3e-4 (a kink of the position warp crosses the tile): the fast path
//                with d* = max | min of the pieces (BP_F_MIN)
//   BP_F_GENERIC two pieces of clearly different slope and/or
//   BP_F_FOUR    rays less than a position apart on the pixel line (the warp compresses them): up to three rays reach a
//                pixel, so four candidate taps are weighted.  Every pixel is located on its own (the generic path).
constexpr int BP_F_GENERIC = 1, BP_F_MIN = 2, BP_F_FOUR = 4, BP_F_KINK = 8;

// What the consumers need to know about one angle of one tile (the producer's plan; 48 bytes)
struct BpPlanRec {
    float4 c4;     // affine piece A of d*: dcol, drow, base - 1/2;  Ah = A0 - B/2
    float4 c5;     // affine piece B
    float2 c6;     // B of piece A, flags (BP_F_*)
    float bB;      // B of piece B
    int aligned;   // first staged slot of the angle's packed sinogram row (a multiple of 4)
};

// Detector coordinate of pixel (line l, position k), in the reference's arithmetic (common.cuh: WarpAngle):
//     d*(l, k) = (Ginv(G(k) - l * delta) - c0_base) / c0_step - 1/2.
// The tile lies inside one stripe of G (tiles are aligned to 64 columns / TH rows, stripes to powers of two; over
// positions [0, 64) G is replaced by its secant, common.cuh: BpWarpF), so z = G(k) - l * delta is affine over the
// tile.  Ginv is piecewise affine in z with kinks where a ray STARTED on a binade boundary; the tile's z-range
// (< 97 wide) usually holds none that matters (2-7 % of the tile x angle pairs at 8192^2 do): d* is then the max (or
// min) of two affine pieces.  The stripe searches run in float on a 96-byte table per angle; only the pieces that
// were picked are evaluated in double.
// (tile origin (row0, col0), real extent tw x th pixels; `padl`: zero detectors in front of a packed sinogram row)
__device__ __forceinline__ BpPlanRec bp_plan_angle(const BpWarpF *__restrict__ bwarp, const BpWarpD *__restrict__ bwarpd, int padl,
                                                   int col0, int row0, int tw, int th, int a) {
    BpPlanRec pl;
    const BpWarpF *wf = bwarp + a;
    const BpWarpD *w = bwarpd + a;
    const bool vert = __ldg(&w->vertical) != 0;
    const int k0 = vert ? col0 : row0, l0 = vert ? row0 : col0;
    const int nk = vert ? tw : th, nl = vert ? th : tw;
    const int ik = k0 < 64 ? 0 : min(WT_NM - 1, 26 - __clz(k0));      // floor(log2 k0) - 5
    const float skf = __ldg(&wf->slope[ik]), iskf = __ldg(&wf->isl[ik]), dlf = __ldg(&wf->delta);
    const float istep = __ldg(&wf->inv_step);
    // z of the tile origin in double (the one quantity that needs it: d* to 1e-5 at |z| ~ 1e4) ...
    const double za = __ldg(&w->cum[ik]) + (double)(k0 - (ik ? (32 << ik) : 0)) * __ldg(&w->slope[ik]) - (double)l0 * __ldg(&w->delta);
    // ... the tile's extent and the stripe searches in float
    const float zaf = (float)za, zk = skf * (float)(nk - 1), zl = -dlf * (float)(nl - 1);
    const float zlo = zaf + fminf(0.0f, zk) + fminf(0.0f, zl), zhi = zaf + fmaxf(0.0f, zk) + fmaxf(0.0f, zl);
    const float zc = 0.5f * (zlo + zhi), azc = fabsf(zc);
    // piece A: the stripe of the tile centre (G's edges lie within 0.1 % of the powers of two, except beyond a
    // stripe where the reference's rays stand still: walk from the power-of-two guess)
    int ia = azc < 64.0f ? 0 : min(WT_NM - 1, (int)((__float_as_uint(azc) >> 23) & 0xffu) - 127 - 5);
    while (ia > 0 && azc < __ldg(&wf->cum[ia])) --ia;
    while (ia + 1 < WT_NM && azc >= __ldg(&wf->cum[ia + 1])) ++ia;
    const bool negc = zc < 0.0f;
    // piece B: across the kink inside (zlo, zhi) that bends d* most -- the edges of A's stripe on the centre's
    // side of zero, and the first edge on the other side if the range straddles zero (a range is < 97 wide)
    int ib = ia;
    bool negb = negc, b_high = false;                                  // b_high: B lies at larger z than the centre
    float worst = 5.0e-5f;                                             // kinks that bend d* by less are ignored
    auto consider = [&](int m, bool neg) {
        if (m < 1 || m >= WT_NM) return;
        const float edge = __ldg(&wf->cum[m]), zb = neg ? -edge : edge;
        const float far_side = zb > zc ? zhi - zb : zb - zlo;          // extent on the side away from the centre
        const float w_ = __ldg(&wf->bend[m]) * far_side;
        if (zb > zlo && zb < zhi && w_ > worst) {
            worst = w_;
            // the stripe on the far side of this kink (further from zero if the kink lies beyond the centre in |z|)
            const bool outward = neg ? (zb < zc) : (zb > zc);
            ib = outward ? m : m - 1;
            negb = neg;
            b_high = zb > zc;
        }
    };
    consider(ia, negc);
    consider(ia + 1, negc);
    if (ia == 0) consider(1, !negc);
    const bool kink = !(ib == ia && negb == negc);
    // affine piece (sgn, i):  Ginv(z) = sgn * pos_i + (z - sgn * gcum_i) / slope_i;  d* at the tile origin in
    // double, then relative to an integer near it in float; the slopes of d* in float
    const double c0_base = __ldg(&w->c0_base), inv_step = __ldg(&w->inv_step);
    auto origin = [&](bool neg, int i) {
        const double pos = i ? (double)(32 << i) : 0.0, cum = __ldg(&w->cum[i]);
        return ((neg ? -pos : pos) + (za - (neg ? -cum : cum)) * __ldg(&w->isl[i]) - c0_base) * inv_step - 0.5;
    };
    const double tA = origin(negc, ia);
    const int ibase = __double2int_rd(tA);
    const float islA = __ldg(&wf->isl[ia]), islB = __ldg(&wf->isl[ib]);
    const float tAf = (float)(tA - (double)ibase), kA = skf * islA * istep, lA = -dlf * islA * istep;
    float tBf = tAf, kB = kA, lB = lA;
    if (kink) {
        tBf = (float)(origin(negb, ib) - (double)ibase);
        kB = skf * islB * istep;
        lB = -dlf * islB * istep;
    }
    // beyond the kink piece B is the right one: d* = max(A, B) if B lies above A there, else min(A, B)
    const float kf = (float)(nk - 1), lf = (float)(nl - 1);
    bool use_max = true;
    if (kink) {      // corner of the tile furthest into B's side in z
        const float ck = (zk > 0.0f) == b_high ? kf : 0.0f, cl = (zl > 0.0f) == b_high ? lf : 0.0f;
        use_max = (tBf + ck * kB + cl * lB) >= (tAf + ck * kA + cl * lA);
    }
    float dmin = 3.0e38f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float ck = (c & 1) ? kf : 0.0f, cl = (c & 2) ? lf : 0.0f;
        const float va = tAf + ck * kA + cl * lA, vb = tBf + ck * kB + cl * lB;
        dmin = fminf(dmin, use_max ? fmaxf(va, vb) : fminf(va, vb));
    }
    // B = len * |dc/dd| at fixed line: the rays' spacing on the pixel line, i.e. their spacing at the start
    // stretched by the warp between the start stripe and the tile's stripe
    const float a0 = __ldg(&wf->a0), bnom = __ldg(&wf->b);
    const float bA = bnom * (__ldg(&wf->slope[ia]) * iskf), bB = kink ? bnom * (__ldg(&wf->slope[ib]) * iskf) : bA;
    // rays closer than a position: a third ray can reach a pixel (weight up to len - B)
    const bool four = fminf(bA, bB) < 0.9985f * a0;
    const bool mild = !kink || fabsf(islA - islB) < 3.0e-4f * islA;
    const bool generic = four || (kink && !mild);
    const int flags = (generic ? BP_F_GENERIC : 0) | (kink && !generic ? BP_F_KINK : 0) | (kink && !use_max ? BP_F_MIN : 0) |
                      (four ? BP_F_FOUR : 0);
    // segment start (one slot of slack below the smallest d*: the shared pairs look one tap down)
    const int lo = ibase + (int)floorf(dmin - 1.0e-3f) - (four ? 2 : 1) + padl;  // >= 0 by construction of padl
    const int aligned = lo & ~3;
    // base is d* - 1/2 relative to the segment start; with f' = frac - 1/2 both tap weights share one
    // constant: w0 = max(0, Ah - B f'), w1 = max(0, Ah + B f'), Ah = A0 - B/2 (= A1 + B/2)
    const float off = (float)(ibase + padl - aligned) - 0.5f;
    pl.c4 = make_float4(vert ? kA : lA, vert ? lA : kA, tAf + off, a0 - 0.5f * bA);
    pl.c5 = make_float4(vert ? kB : lB, vert ? lB : kB, tBf + off, a0 - 0.5f * bB);
    pl.c6 = make_float2(bA, __int_as_float(flags));
    pl.bB = bB;
    pl.aligned = aligned;
    return pl;
}

// Plans of a whole small geometry, computed once at geometry creation (the same function the kernel runs inline):
// table[(tile_y * ntx + tile_x) * A + a].  For small problems the producer's plan -- ~370 dependent instructions with
// four or five dependent table loads -- is the critical path of every pipeline stage.
template <int TH>
__global__ void bp_plan_table_kernel(const BpWarpF *bwarp, const BpWarpD *bwarpd, int padl, int R, int C, int A, BpPlanRec *table) {
    const int a = blockIdx.z * blockDim.x + threadIdx.x;
    if (a >= A) return;
    const int col0 = blockIdx.x * BP_TW, row0 = blockIdx.y * TH;
    table[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * A + a] =
        bp_plan_angle(bwarp, bwarpd, padl, col0, row0, min(BP_TW, C - col0), min(TH, R - row0), a);
}


// Epilogues with kHasSums accumulate up to BP_MAX_SUMS per-thread sums (operator() gets a float* to them); the kernel
// reduces them over the warp and hands them to epi.finish<W>(p, g, slot, nslots, sums): slot = (tile, warp), so the
// partial sums are written in a fixed place and reduced in a fixed order by a later kernel (deterministic).
constexpr int BP_MAX_SUMS = 4;

// Epilogue contract: epi.operator()<W>(p, g, row, col, val[W]) handles the W images of group g for one pixel.  Under
// MIRROR the V components are V/2 images x {pixel (row, col), its point reflection (R-1-row, C-1-col)}.
// Resident CTAs per SM: three for the float4 kernels (72 registers: 32 accumulators + 12 live taps).  Measured for the
// float / float2 kernels (BP at 2048^2 x 1800 / 512^2 x 360, ms): one image 3 / 4 / 5 / 6 CTAs: 3.39 / 3.41 / 3.64 (spills) /
// 4.26; two images 3 / 4 CTAs: 4.00 / 3.99 and 0.082 / 0.079.
#ifndef WT_BP_MINCTAS_V1
#define WT_BP_MINCTAS_V1 3
#endif
#ifndef WT_BP_MINCTAS_V2
#define WT_BP_MINCTAS_V2 4
#endif
template <int V> struct BpMinCtas { static constexpr int value = V == 1 ? WT_BP_MINCTAS_V1 : V == 2 ? WT_BP_MINCTAS_V2 : 3; };

template <int V, class Epi, bool MIRROR, int TH>
__global__ void __launch_bounds__(BP_BLOCK, BpMinCtas<V>::value) bp_kernel(BpParams p, Epi epi) {
    typedef typename VecOf<V>::T VecT;
    constexpr int BP_TH = TH, BP_RPT = BpTile<TH>::RPT, BP_SEGW = BpTile<TH>::SEGW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BpSmem<V, TH> &sm = *reinterpret_cast<BpSmem<V, TH> *>(smem_raw);

    // Small volumes (64 x 16 tiles, fewer tiles than SMs): the CTAs of a cluster share a tile, each backprojects a
    // range of the angle blocks, and rank 0 adds the partial sums in rank order (through distributed shared memory)
    // before the epilogue -- deterministic, and the epilogue still sees the complete sum.
    constexpr bool SPLIT = TH == 16;
    const int cs = SPLIT ? p.csplit : 1;
    const int crank = cs > 1 ? (int)(blockIdx.z % (unsigned)cs) : 0;
    const int g = cs > 1 ? (int)(blockIdx.z / (unsigned)cs) : (int)blockIdx.z;
    int a_begin = p.a_begin, a_end = p.a_end;
    if (cs > 1) {
        const int per = ((a_end - a_begin + BP_AB - 1) / BP_AB + cs - 1) / cs * BP_AB;    // angles per rank, whole blocks
        a_begin = min(a_end, a_begin + crank * per);
        a_end = min(a_end, a_begin + per);
    }
    const int col0 = blockIdx.x * BP_TW, row0 = p.row_begin + blockIdx.y * BP_TH;
    const int tw = min(BP_TW, p.C - col0), th = min(BP_TH, p.row_end - row0);  // real pixels of this tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nblk = (a_end - a_begin + BP_AB - 1) / BP_AB;

    WT_TRACE_POINT(1, 0, threadIdx.x == 0);
    WT_TRACE_POINT(1, 1, threadIdx.x == 0);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < BP_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], BP_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Producer warp: lane = one angle of the block.  plan() computes the angle's constants for this tile into registers
    // -- BEFORE the producer waits for the stage to be released, so the arithmetic is off the refill path -- and
    // issue() publishes them and starts the TMA copy of the angle's sinogram segment.
    //
    struct Plan { BpPlanRec r; const VecT *src; };
    // (from the geometry's precomputed table when it has one -- small geometries, whole tiles of the unmirrored volume)
    const bool tabled = p.plans != nullptr && p.Rproc == p.R && (p.row_end == p.R || p.row_end % BP_TH == 0);
    const BpPlanRec *trow = p.plans + ((size_t)(row0 / BP_TH) * ((p.C + BP_TW - 1) / BP_TW) + blockIdx.x) * p.A;
    auto plan = [&](int blk) {
        Plan pl;
        pl.src = nullptr;
        const int a = a_begin + blk * BP_AB + lane;
        const int nvalid = min(BP_AB, a_end - (a_begin + blk * BP_AB));
        if (lane < nvalid) {
            if (tabled) {
                const float4 *q = reinterpret_cast<const float4 *>(trow + a);
                const float4 q2 = __ldg(q + 2);
                pl.r.c4 = __ldg(q);
                pl.r.c5 = __ldg(q + 1);
                pl.r.c6 = make_float2(q2.x, q2.y);
                pl.r.bB = q2.z;
                pl.r.aligned = __float_as_int(q2.w);
            } else {
                pl.r = bp_plan_angle(p.bwarp, p.bwarpd, p.padl, col0, row0, tw, th, a);
            }
            pl.src = reinterpret_cast<const VecT *>(p.sp) + ((size_t)g * p.A + a) * p.pitch + pl.r.aligned;
        }
        return pl;
    };
    auto issue = [&](int blk, const Plan &pl) {
        const int s = blk % BP_STAGES;
        const int nvalid = min(BP_AB, a_end - (a_begin + blk * BP_AB));
        if (lane < nvalid) {
            sm.cst[s][lane] = pl.r.c4;
            sm.cst2[s][lane] = pl.r.c5;
            sm.cst3[s][lane] = pl.r.c6;
            sm.cst4[s][lane] = pl.r.bB;
        }
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();  // consumers' generic-proxy reads of this stage precede the refill
            mbar_arrive_expect_tx(&sm.full[s], (uint32_t)(nvalid * BP_SEGW * sizeof(VecT)));
        }
        __syncwarp();
        // (per-lane copies compile to an elect / R2UR loop, ~150 cycles per copy; one lane issuing the 16 copies from
        // shuffled addresses compiles to 16 such one-trip loops and measured no faster: 23.29 vs 23.15 ms at C4, 39.0 vs
        // 38.7 us per 256^2 SIRT iteration)
        if (lane < nvalid) tma_load_1d(&sm.seg[s][lane][0], pl.src, (uint32_t)(BP_SEGW * sizeof(VecT)), &sm.full[s]);
    };

    if (warp == BP_THREADS / 32) {
        // runs ahead of the consumers by up to BP_STAGES angle blocks; consumer warps never wait for each other
        for (int blk = 0; blk < nblk; ++blk) {
            const Plan pl = plan(blk);
            if (blk >= BP_STAGES) mbar_wait(&sm.empty[blk % BP_STAGES], (uint32_t)(((blk / BP_STAGES) - 1) & 1));
            issue(blk, pl);
        }
        if (cs > 1) cluster_sync_all();     // (the consumers' one, below)
        return;
    }
    WT_TRACE_POINT(1, 2, threadIdx.x == 0);

    // this thread's pixels: columns lane, lane+32; rows in vertically adjacent pairs 2*warp + 16*j + {0, 1} (clamped
    // into the real tile so that every smem index stays inside the staged segment; clamped results are dropped)
    float jc[2], ir[BP_RPT];
#pragma unroll
    for (int c = 0; c < 2; ++c) jc[c] = (float)min(lane + 32 * c, tw - 1);
#pragma unroll
    for (int r = 0; r < BP_RPT; ++r) ir[r] = (float)min(bp_row_of(warp, r), th - 1);

    constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23
    constexpr int MAGIC_BITS = 0x4B400000;
    float acc[BP_RPT][2][V];
#pragma unroll
    for (int r = 0; r < BP_RPT; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[r][c][i] = 0.0f;

    // One pixel: cell and fraction from the magic-number rounding, two taps, V FMAs each.
    auto locate = [&](float ds, int &idx, float &f) {
        const float t = ds + MAGIC;            // round(d* - 1/2) = floor(d*) (ties: the extra tap has weight 0)
        f = ds - (t - MAGIC);                  // frac(d*) - 1/2
        idx = __float_as_int(t) - MAGIC_BITS;
    };
    // (`seg` minus MAGIC_BITS elements, indexed with the raw bits of the rounded sum: one LEA per tap address)
    auto single = [&](float ds, const float4 &c4, float bw, const VecT *seg, float (&a)[V]) {
        const float t = ds + MAGIC;            // round(d* - 1/2) = floor(d*) (ties: the extra tap has weight 0)
        const float f = ds - (t - MAGIC);      // frac(d*) - 1/2
        check_index(__float_as_int(t) - MAGIC_BITS, __float_as_int(t) - MAGIC_BITS + 1, BP_SEGW);
        const VecT *tap = (seg - MAGIC_BITS) + __float_as_int(t);
        const float w0 = fmaxf(0.0f, fmaf(-bw, f, c4.w)), w1 = fmaxf(0.0f, fmaf(bw, f, c4.w));
        float v0[V], v1[V];
        unpack<V>(tap[0], v0);
        unpack<V>(tap[1], v1);
        tap_pair_fma<V>(w0, v0, w1, v1, a);
    };
    // Two pixels at once (float and float2 kernels, which are bound by the instruction issue rate: ~14 issue slots per
    // pixel and angle): cell, fraction and weights of both go through the packed fp32 pipe (FADD2 / FFMA2, each half
    // rounded like the scalar operation; x - y is written fma(y, -1, x), whose product is exact), and with one image per
    // position so do the tap FMAs -- the bits of `single`, ~9 slots per pixel.
    auto pixel_pair = [&](float2 ds, const float4 &c4, float bw, const VecT *seg, float (&a0)[V], float (&a1)[V]) {
        const float2 t = __fadd2_rn(ds, make_float2(MAGIC, MAGIC));
        const float2 u = __fadd2_rn(t, make_float2(-MAGIC, -MAGIC));
        const float2 f = __ffma2_rn(u, make_float2(-1.0f, -1.0f), ds);      // frac(d*) - 1/2
        check_index(__float_as_int(t.x) - MAGIC_BITS, __float_as_int(t.x) - MAGIC_BITS + 1, BP_SEGW);
        check_index(__float_as_int(t.y) - MAGIC_BITS, __float_as_int(t.y) - MAGIC_BITS + 1, BP_SEGW);
        const VecT *ta = (seg - MAGIC_BITS) + __float_as_int(t.x), *tb = (seg - MAGIC_BITS) + __float_as_int(t.y);
        float2 w0 = __ffma2_rn(f, make_float2(-bw, -bw), make_float2(c4.w, c4.w));
        float2 w1 = __ffma2_rn(f, make_float2(bw, bw), make_float2(c4.w, c4.w));
        w0.x = fmaxf(0.0f, w0.x); w0.y = fmaxf(0.0f, w0.y);
        w1.x = fmaxf(0.0f, w1.x); w1.y = fmaxf(0.0f, w1.y);
        float va0[V], va1[V], vb0[V], vb1[V];
        unpack<V>(ta[0], va0);
        unpack<V>(ta[1], va1);
        unpack<V>(tb[0], vb0);
        unpack<V>(tb[1], vb1);
        if constexpr (V == 1) {
            float2 s2 = make_float2(a0[0], a1[0]);
            s2 = __ffma2_rn(w0, make_float2(va0[0], vb0[0]), __ffma2_rn(w1, make_float2(va1[0], vb1[0]), s2));
            a0[0] = s2.x;
            a1[0] = s2.y;
        } else {
            tap_pair_fma<V>(w0.x, va0, w1.x, va1, a0);
            tap_pair_fma<V>(w0.y, vb0, w1.y, vb1, a1);
        }
    };
    // A vertically adjacent pixel pair with shared taps (float4 kernels).  The upper pixel is located as `single`
    // does; the lower one sits drow (|drow| <= 1) detectors further, so its cell is the same one or the next one in
    // the direction of drow and its fraction follows with one add and one compare -- no second round trip through the
    // magic number, no dependent address.  The four taps of the pair lie on three consecutive detectors: 3 LDS
    // instead of 4 (shared-memory wavefronts are what bounds this kernel).  The lower pixel puts three weights on the
    // three taps, (w0, w1, 0) or (0, w0, w1) depending on its cell: 3 selects on scalars instead of 2 V selects on
    // taps, and fma(0, tap, acc) leaves acc untouched, so the bits are those of the two-tap form.
    // POS = sign of drow, uniform per angle.
    auto shared_pair = [&](auto pos_tag, float ds, const float4 &c4, float bw, const VecT *seg, float (&a0)[V],
                           float (&a1)[V]) {
        constexpr bool POS = decltype(pos_tag)::value;
        int idx;
        float f0;
        locate(ds, idx, f0);
        check_index(POS ? idx : idx - 1, POS ? idx + 2 : idx + 1, BP_SEGW);
        const float g = f0 + c4.y;
        const bool moved = POS ? (g >= 0.5f) : (g < -0.5f);
        const float f1 = moved ? (POS ? g - 1.0f : g + 1.0f) : g;
        const float w00 = fmaxf(0.0f, fmaf(-bw, f0, c4.w)), w01 = fmaxf(0.0f, fmaf(bw, f0, c4.w));
        const float w10 = fmaxf(0.0f, fmaf(-bw, f1, c4.w)), w11 = fmaxf(0.0f, fmaf(bw, f1, c4.w));
        float v0[V], v1[V], vx[V];
        unpack<V>(seg[idx], v0);
        unpack<V>(seg[idx + 1], v1);
        unpack<V>(seg[POS ? idx + 2 : idx - 1], vx);
        tap_pair_fma<V>(w00, v0, w01, v1, a0);
        // taps in detector order: (v0, v1, vx) for POS, (vx, v0, v1) otherwise
        const float wa = POS ? (moved ? 0.0f : w10) : (moved ? w10 : 0.0f);
        const float wb = POS ? (moved ? w10 : w11) : (moved ? w11 : w10);
        const float wc = POS ? (moved ? w11 : 0.0f) : (moved ? 0.0f : w11);
        if constexpr (POS) tap_triple_fma<V>(wa, v0, wb, v1, wc, vx, a1);
        else tap_triple_fma<V>(wa, vx, wb, v0, wc, v1, a1);
    };
    // KINK: d* = max | min of two affine pieces of nearly equal slope (the lower pixel of a pair and the tap weights
    // take piece A's slope: < 3e-4 detector off)
    auto one_angle = [&](auto pos_tag, auto kink_tag, const float4 &c4, const float4 &c5, bool use_min, float bw, const VecT *seg) {
        constexpr bool KINK = decltype(kink_tag)::value;
        auto dstar = [&](float j, float i) {
            const float a = fmaf(j, c4.x, fmaf(i, c4.y, c4.z));
            if (!KINK) return a;
            const float b = fmaf(j, c5.x, fmaf(i, c5.y, c5.z));
            return use_min ? fminf(a, b) : fmaxf(a, b);
        };
        if constexpr (!KINK && BpShare<V>::value == 0 && WT_BP_PAIRPIX) {
#pragma unroll
            for (int j = 0; j < BP_RPT / 2; ++j) {
                const float2 dr2 = __ffma2_rn(make_float2(ir[2 * j], ir[2 * j + 1]), make_float2(c4.y, c4.y), make_float2(c4.z, c4.z));
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    pixel_pair(__ffma2_rn(make_float2(jc[c], jc[c]), make_float2(c4.x, c4.x), dr2), c4, bw, seg, acc[2 * j][c], acc[2 * j + 1][c]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < BP_RPT / 2; ++j) {
                const float dr0 = fmaf(ir[2 * j], c4.y, c4.z), dr0b = KINK ? fmaf(ir[2 * j], c5.y, c5.z) : 0.0f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float da = fmaf(jc[c], c4.x, dr0), db = KINK ? fmaf(jc[c], c5.x, dr0b) : 0.0f;
                    const float ds0 = !KINK ? da : (use_min ? fminf(da, db) : fmaxf(da, db));
                    if (2 * j + c < BpShare<V>::value) {
                        shared_pair(pos_tag, ds0, c4, bw, seg, acc[2 * j][c], acc[2 * j + 1][c]);
                    } else {
                        single(ds0, c4, bw, seg, acc[2 * j][c]);
                        single(dstar(jc[c], ir[2 * j + 1]), c4, bw, seg, acc[2 * j + 1][c]);
                    }
                }
            }
        }
    };
    // Generic path (BP_F_GENERIC): d* = max or min of the two affine pieces, each pixel located on its own with the tap
    // spacing of the piece it falls on; with BP_F_FOUR the two outer candidates idx - 1 and idx + 2 get their weights
    // max(0, len - B * distance) too (they vanish unless the rays are closer than a position).
    auto one_angle_generic = [&](const float4 &c4, const float4 &c5, float bwa, float bwb, int flags, const VecT *seg) {
        const bool use_min = (flags & BP_F_MIN) != 0, four = (flags & BP_F_FOUR) != 0;
#pragma unroll
        for (int r = 0; r < BP_RPT; ++r) {
            const float dra = fmaf(ir[r], c4.y, c4.z), drb = fmaf(ir[r], c5.y, c5.z);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float da = fmaf(jc[c], c4.x, dra), db = fmaf(jc[c], c5.x, drb);
                const bool on_b = use_min ? (db < da) : (db > da);
                const float ds = on_b ? db : da, bw = on_b ? bwb : bwa, ah = on_b ? c5.w : c4.w;
                int idx;
                float f;
                locate(ds, idx, f);
                idx = min(max(idx, 1), BP_SEGW - 3);       // a tile whose shadow outgrows the segment (rays compressed by > 10 %)
                const float w0 = fmaxf(0.0f, fmaf(-bw, f, ah)), w1 = fmaxf(0.0f, fmaf(bw, f, ah));
                float v0[V], v1[V];
                unpack<V>(seg[idx], v0);
                unpack<V>(seg[idx + 1], v1);
                tap_pair_fma<V>(w0, v0, w1, v1, acc[r][c]);
                if (four) {
                    const float wm = fmaxf(0.0f, fmaf(-bw, f, ah - bw)), wp = fmaxf(0.0f, fmaf(bw, f, ah - bw));
                    unpack<V>(seg[idx - 1], v0);
                    unpack<V>(seg[idx + 2], v1);
                    tap_pair_fma<V>(wm, v0, wp, v1, acc[r][c]);
                }
            }
        }
    };

    for (int blk = 0; blk < nblk; ++blk) {
        const int s = blk % BP_STAGES;
        mbar_wait(&sm.full[s], (uint32_t)((blk / BP_STAGES) & 1));
        WT_TRACE_POINT(1, 3, threadIdx.x == 0 && blk == 0);
        const int nvalid = min(BP_AB, a_end - (a_begin + blk * BP_AB));
        for (int k = 0; k < nvalid; ++k) {
            const float4 c4 = sm.cst[s][k];
            const float2 c6 = sm.cst3[s][k];
            const VecT *seg = &sm.seg[s][k][0];
            const int flags = __float_as_int(c6.y);
            if (flags == 0) {       // the common case first
                if (BpShare<V>::value == 0 || c4.y >= 0.0f) one_angle(std::true_type(), std::false_type(), c4, c4, false, c6.x, seg);
                else one_angle(std::false_type(), std::false_type(), c4, c4, false, c6.x, seg);
            } else if (flags & BP_F_GENERIC) one_angle_generic(c4, sm.cst2[s][k], c6.x, sm.cst4[s][k], flags, seg);
            else if (flags & BP_F_KINK) {
                const float4 c5 = sm.cst2[s][k];
                const bool use_min = (flags & BP_F_MIN) != 0;
                if (BpShare<V>::value == 0 || c4.y >= 0.0f) one_angle(std::true_type(), std::true_type(), c4, c5, use_min, c6.x, seg);
                else one_angle(std::false_type(), std::true_type(), c4, c5, use_min, c6.x, seg);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);  // this warp is done with stage s
    }

    WT_TRACE_POINT(1, 4, threadIdx.x == 0);
    if (cs > 1) {
        // partial sums of ranks 1 .. cs-1 travel to rank 0's shared memory: part[rank - 1][accumulator][thread]
        float *part = reinterpret_cast<float *>(smem_raw + sizeof(BpSmem<V, TH>));
        constexpr int NACC = BP_RPT * 2 * V;
        if (crank > 0) {
#pragma unroll
            for (int r = 0; r < BP_RPT; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        st_cluster_f32(part + ((size_t)(crank - 1) * NACC + (r * 2 + c) * V + i) * BP_THREADS + threadIdx.x, 0, acc[r][c][i]);
        }
        cluster_sync_all();
        WT_TRACE_POINT(1, 5, threadIdx.x == 0);
        if (crank > 0) return;
        for (int k = 1; k < cs; ++k)
#pragma unroll
            for (int r = 0; r < BP_RPT; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        acc[r][c][i] += part[((size_t)(k - 1) * NACC + (r * 2 + c) * V + i) * BP_THREADS + threadIdx.x];
    }

    float sums[BP_MAX_SUMS];
#pragma unroll
    for (int q = 0; q < BP_MAX_SUMS; ++q) sums[q] = 0.0f;
#pragma unroll
    for (int r = 0; r < BP_RPT; ++r) {
        const int row = row0 + bp_row_of(warp, r);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = col0 + lane + 32 * c;
            if (row < p.row_end && col < p.C) {
                if constexpr (MIRROR) {
                    static_assert(!Epi::kHasSums, "epilogues with sums are not mirror-paired");
                    constexpr int W = V / 2;
                    float direct[W], mirrored[W];
#pragma unroll
                    for (int i = 0; i < W; ++i) { direct[i] = acc[r][c][i]; mirrored[i] = acc[r][c][W + i]; }
                    epi.template operator()<W>(p, g, row, col, direct);
                    const int mr = p.R - 1 - row;     // the middle row of an odd-height image is its own mirror row
                    if (mr != row) epi.template operator()<W>(p, g, mr, p.C - 1 - col, mirrored);
                } else if constexpr (Epi::kHasSums) {
                    epi.template operator()<V>(p, g, row, col, acc[r][c], sums);
                } else {
                    epi.template operator()<V>(p, g, row, col, acc[r][c]);
                }
            }
        }
    }
    WT_TRACE_POINT(1, 6, threadIdx.x == 0);
    WT_TRACE_POINT(1, 7, threadIdx.x == 0);
    if constexpr (Epi::kHasSums && !MIRROR) {
#pragma unroll
        for (int q = 0; q < BP_MAX_SUMS; ++q) sums[q] = warp_sum(sums[q]);
        if (lane == 0)
            epi.template finish<V>(p, g, (int)((blockIdx.y * gridDim.x + blockIdx.x) * (BP_THREADS / 32) + warp),
                                   (int)(gridDim.x * gridDim.y * (BP_THREADS / 32)), sums);
    }
}

// Tile height.  A 64 x 32 tile is the faster one when there are CTAs to spare (23.2 vs 24.1 ms at 2048^2 x 16 images),
// but a 512^2 volume has only 128 of them for 148 SMs: volumes of at most one 64 x 32 tile per SM use 64 x 16 tiles
// (measured BP: 256^2 43 -> 30 us, 512^2 78 -> 66 us per image).  The choice depends on the volume alone, never on the
// number of images, so that an image's bits do not depend on the batch it travels in.
inline int bp_tile_height(int rows, int cols) {
    const long tiles32 = (long)((cols + BP_TW - 1) / BP_TW) * ((rows + 31) / 32);
    return tiles32 <= 148 ? 16 : 32;
}

// Cluster split of the angle blocks (bp_kernel): only for volumes whose 64 x 16 tiles leave SMs idle (up to four CTAs
// per SM after the split), as long as every rank keeps at least two angle blocks -- measured on the C-side SIRT loop:
// 256^2 75 -> 39 us (4 ranks), 384^2 130 -> 95 us, 512^2 191 -> 161 us (2 ranks) per iteration.  Like the tile height it
// depends on the geometry alone, never on the number of images or on the epilogue, so an image's bits do not depend on
// the call it travels in.
constexpr int BP_MAX_SPLIT = 8;
inline int bp_cluster_split(int rows, int cols, int nangles) {
    if (bp_tile_height(rows, cols) != 16) return 1;
    static const int off = getenv("WT_NO_SPLIT") ? 1 : 0;       // A/B measurements
    if (off) return 1;
    static const int forced = getenv("WT_BP_SPLIT") ? atoi(getenv("WT_BP_SPLIT")) : 0;
    if (forced >= 1 && forced <= BP_MAX_SPLIT) return forced;
    const long tiles = (long)((cols + BP_TW - 1) / BP_TW) * ((rows + 15) / 16);
    const int nblk = (nangles + BP_AB - 1) / BP_AB;
    int cs = 1;
    while (cs < 4 && tiles * cs * 2 <= 4 * 148 && nblk >= 4 * cs) cs *= 2;
    return cs;
}

template <int V, class Epi, bool MIRROR, int TH>
inline cudaError_t launch_bp_th(const BpParams &p, const Epi &epi, int groups, cudaStream_t st) {
    // cluster split (64 x 16 tiles only): rank 0 receives the other ranks' accumulators behind the pipeline buffers
    constexpr size_t part_bytes = (size_t)BpTile<TH>::RPT * 2 * V * BP_THREADS * sizeof(float);
    const int cs = TH == 16 ? bp_cluster_split(p.R, p.C, p.A) : 1;
    const size_t smem_max = sizeof(BpSmem<V, TH>) + (TH == 16 ? (BP_MAX_SPLIT - 1) * part_bytes : 0);
    const size_t smem = sizeof(BpSmem<V, TH>) + (cs - 1) * part_bytes;
    static bool configured[64] = {};   // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(bp_kernel<V, Epi, MIRROR, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    BpParams q = p;
    q.csplit = cs;
    dim3 grid((p.C + BP_TW - 1) / BP_TW, (p.row_end - p.row_begin + TH - 1) / TH, groups * cs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(BP_BLOCK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    LaunchAttrs la;
    if (cs > 1) la.cluster(1, 1, cs);
    cfg.attrs = la.at;
    cfg.numAttrs = la.n;
    return cudaLaunchKernelEx(&cfg, bp_kernel<V, Epi, MIRROR, TH>, q, epi);
}

template <int V, class Epi, bool MIRROR = false>
inline cudaError_t launch_bp(const BpParams &p, const Epi &epi, int groups, cudaStream_t st) {
    if (bp_tile_height(p.R, p.C) == 16) return launch_bp_th<V, Epi, MIRROR, 16>(p, epi, groups, st);
    return launch_bp_th<V, Epi, MIRROR, 32>(p, epi, groups, st);
}

}  // namespace wt
