// K1: ray-driven Joseph forward projector (ASTRA 'linear' projector semantics),
// V images interleaved per position so that one 16-byte shared-memory load feeds V updates.
//
// Work decomposition.  Thread <-> ray (angle a, detector d).  A CTA owns FP_DT detectors x FP_AT angles
// of one "angle tile" (consecutive angles of the same orientation whose rays stay within one window,
// grouped on the host at geometry creation) and marches over the lines of the volume in chunks of
// FP_LB lines.  For every chunk one warp asks the TMA unit (cp.async.bulk, one 1-D bulk copy per line)
// to stage the FP_W-position window of those lines that the tile's rays cross; completion is tracked
// with one mbarrier per stage, FP_STAGES deep, so the copies of later chunks overlap the gathers of the
// current one.  "Vertical" angles read the row-major packed volume, "horizontal" angles the transposed
// packed copy -- the same loop either way.
//
// Lane mapping.  A quarter-warp (the unit the LSU serves per shared-memory wavefront for 16-byte loads)
// holds 4 adjacent detectors x 2 adjacent angles: their two Joseph taps fall within < 8 consecutive
// 16-byte positions, i.e. distinct banks or the same word -> conflict-free LDS.128 at the full
// 128 B/clk/SM.  (The first version of this kernel gathered through L1 with __ldg: a warp's 32 adjacent
// detectors span 5-7 cache lines per load and ncu showed l1tex at 98% with 2.3x the minimum wavefronts;
// see profiles/r1_ncu_full_v1_summary.txt.)
//
// Positions.  WT_POS_ASTRA (default): every thread repeats the reference's float32 arithmetic literally -- the
// position at line 0 with ASTRA's own sequence of float32 operations, then one float32 addition per line from line 0
// on (oracle/astra_linear.c:88-93,119-120), including the lines before the ray enters the volume (a serial chain of
// FADDs in the prologue: 0.26 N^2 A additions against 11.7 N^2 A instructions of gather work).  The staging windows
// and line ranges come from the same sum in its warped closed form (common.cuh: WarpAngle).  WT_POS_EXACT and the
// mirror-paired kernels evaluate c_0 + l * delta per chunk (float64 bases).
#pragma once
#include "common.cuh"

namespace wt {

constexpr int FP_DT = 32;       // detectors per CTA
constexpr int FP_AT = 8;        // angles per CTA (one angle tile)
constexpr int FP_THREADS = FP_DT * FP_AT;  // consumer threads (one ray each)
constexpr int FP_BLOCK = FP_THREADS + 32;  // + one producer warp that only drives the TMA pipeline
#ifndef WT_FP_LB
#define WT_FP_LB 16
#endif
#ifndef WT_FP_STAGES
#define WT_FP_STAGES 2
#endif
constexpr int FP_LB = WT_FP_LB;  // lines per staged chunk
constexpr int FP_W = 96;        // positions per staged line window
constexpr int FP_STAGES = WT_FP_STAGES;   // float4 kernels (and the least any kernel has: epilogues size their scratch with it)
// The float / float2 kernels get through a chunk 3-4 times faster than the float4 kernel and their stages are 6 / 12 KB,
// so a deeper pipeline costs no occupancy: measured at 2048^2 x 1800, 2 / 3 / 4 / 6 / 8 stages: one image 3.71 / 3.66 /
// 3.65 / 3.62 / 3.62 ms, two images 4.31 / 4.16 / - / 4.23 / 4.23 ms (512^2, two images: 8 stages cost occupancy, 0.111 vs 0.086).
#ifndef WT_FP_STAGES_V1
#define WT_FP_STAGES_V1 4
#endif
#ifndef WT_FP_STAGES_V2
#define WT_FP_STAGES_V2 3
#endif
template <int V> struct FpStages { static constexpr int value = V == 1 ? WT_FP_STAGES_V1 : V == 2 ? WT_FP_STAGES_V2 : FP_STAGES; };
// budget of the window for the host-side tiling: spread of ray positions across the tile must leave room
// for the rounding margin (2 left, 3 right) and the 16-byte alignment of the window start (<= 3 positions)
// (the 9th position of the margin absorbs the residual of the warped closed form that places the windows)
constexpr int FP_WFIT = FP_W - 9;
constexpr int FP_MAX_PHASES = 4;
// Staging copies of a chunk: ONE tiled TMA copy (cp.async.bulk.tensor, a box of FP_LB lines x FP_W positions of the
// packed copy) instead of FP_LB one-line bulk copies.  A per-lane cp.async.bulk compiles to a vote / R2UR loop that
// issues the lanes' copies one after the other at ~150 cycles each: 1.2 us per chunk, which is all of the kernel at
// 256^2 (time proportional to the number of copies, whatever their size or the pipeline depth) and, with four CTAs per
// SM, within 25 % of the chunk time at 2048^2.  WT_FP_TENSOR_TMA=0 keeps the one-line copies.
#ifndef WT_FP_TENSOR_TMA
#define WT_FP_TENSOR_TMA 1
#endif

__host__ __device__ inline int fp_pitch(int npos) {
    int p = npos + 2 * VPAD;
    if (p < FP_W) p = FP_W;
    return (p + 3) & ~3;
}

// ---- packing: plain (n, R, C) -> VN (G, R, pitch(C), V) and VT (G, C, pitch(R), V), zero padded -------
// position k of a line lives at slot k + VPAD; everything else in the line is zero.
// MIRROR: the group holds V0 = V/2 images; components [V0, V) are their point reflections x[R-1-r][C-1-c].  A
// parallel-beam ray (a, d) through the reflected image is the ray (a, D-1-d) through the image, so one or two images
// run through the V = 2 / V = 4 kernels with half the rays and shared index arithmetic (DESIGN.md "mirror pairing").
template <int V, bool MIRROR = false>
__global__ void __launch_bounds__(256) pack_vol_kernel(const float *__restrict__ vol, float *__restrict__ vn,
                                                       float *__restrict__ vt, int R, int C, int nimg, int copies) {
    static_assert(!MIRROR || V % 2 == 0, "mirror pairing needs an even interleave width");
    typedef typename VecOf<V>::T VecT;
    constexpr int V0 = MIRROR ? V / 2 : V;   // images per group
    __shared__ float tile[V][32][33];
    const int g = blockIdx.z;
    const int r0 = blockIdx.y * 32 - VPAD, c0 = blockIdx.x * 32 - VPAD;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Cp = fp_pitch(C), Rp = fp_pitch(R);
    const size_t img = (size_t)R * C;
    const float *src = vol + (size_t)g * V0 * img;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rr = ty + 8 * k;
        const int r = r0 + rr, c = c0 + tx;
        float v[V];
        const bool in = (r >= 0 && r < R && c >= 0 && c < C);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            // components past the last image (padding of the last group) are zero
            const bool real = g * V0 + (MIRROR && i >= V0 ? i - V0 : i) < nimg;
            if (MIRROR && i >= V0) v[i] = in && real ? __ldg(src + (i - V0) * img + (size_t)(R - 1 - r) * C + (C - 1 - c)) : 0.0f;
            else v[i] = in && real ? __ldg(src + i * img + (size_t)r * C + c) : 0.0f;
            tile[i][rr][tx] = v[i];
        }
        if ((copies & 1) && r >= 0 && r < R && c + VPAD < Cp)
            reinterpret_cast<VecT *>(vn)[((size_t)g * R + r) * Cp + (c + VPAD)] = pack<V>(v);
    }
    if (!(copies & 2)) return;      // uniform over the grid: no transposed copy wanted
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cc = ty + 8 * k;
        const int c = c0 + cc, r = r0 + tx;
        if (c >= 0 && c < C && r + VPAD < Rp) {
            float v[V];
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = tile[i][tx][cc];
            reinterpret_cast<VecT *>(vt)[((size_t)g * C + c) * Rp + (r + VPAD)] = pack<V>(v);
        }
    }
}

// copies: bit 0 = row-major copy VN (read by vertical angles), bit 1 = transposed copy VT (horizontal angles)
template <int V, bool MIRROR = false>
inline void launch_pack_vol(const float *vol, float *vn, float *vt, int R, int C, int groups, int nimg, int copies, cudaStream_t st) {
    // the tile grid covers the padded extents of both copies (pitch may exceed npos + 2*VPAD for tiny volumes)
    const int ext_c = max(fp_pitch(C), C + 2 * VPAD), ext_r = max(fp_pitch(R), R + 2 * VPAD);
    dim3 grid((ext_c + 31) / 32, (ext_r + 31) / 32, groups);
    pack_vol_kernel<V, MIRROR><<<grid, dim3(32, 8), 0, st>>>(vol, vn, vt, R, C, nimg, copies);
}

// ---- projector ------------------------------------------------------------------------------------
struct FpTile {
    int first, count;  // angles [first, first + count), count <= FP_AT, all of one orientation
};

struct FpParams {
    const float *vn;
    const float *vt;
    const FpAngle *ang;
    const FpTile *tiles;
    int ntiles;
    int R, C, D, A;
    int nimg;          // images of this launch; the last group is padded with zero components up to V
    int Dproc;         // detectors that get a thread: D, or ceil(D/2) under mirror pairing
    int a_begin, a_end;
    const WarpAngleF *warpf; // per angle: float copy of the position warp (identity under WT_POS_EXACT)
    int astra;               // 1: positions are the reference's float32 running sums (WT_POS_ASTRA)
    float Ex, Ey;            // -C/2 + 1/2, R/2 - 1/2 as float32 (ASTRA's volume corner)
    // CTA launch order (see fp_block): tiles grouped into phases (one per orientation)
    const int *order;  // tile indices, phase by phase
    int nphase;
    int phase_start[FP_MAX_PHASES + 1];  // prefix offsets into `order`
    int ndblk;         // detector blocks per tile
    int csplit;        // CTAs of a cluster that share a (tile, detector block), each taking a range of the line chunks (1: none)
};

// Launch order.  The grid is linear in x: phase by phase (vertical angles read VN, horizontal angles VT), inside a
// phase detector block by detector block, inside a detector block all the phase's angle tiles.  The rays of one
// detector block over all angles of a phase cover a "bow tie" of the packed volume copy (at most ~60 % of it, 40 MB
// of a float4 group's 67 MB at 2048^2) that slides by FP_DT positions per detector block, so the lines the resident
// CTAs read stay in L2 until the bow tie has moved past them and every line comes from DRAM about once per phase.  Measured (ncu, one float4 group at
// 2048^2 x 1800): 156 MB read for 134 MB of packed copies with one phase per orientation; tile-major order -- every
// wave of CTAs sweeping a whole copy that exceeds the ~60 MB effective L2 share -- read 2.3 GB, 0.81 GB with the sweep
// direction alternating per wave; four phases (orientation x slope sign) read each copy twice, 200 MB.
// Logical indices, not blockIdx, name a CTA's partial-sum slot, so results do not depend on the launch order.
struct FpBlock { int tile, dblk; };
__device__ __forceinline__ FpBlock fp_block(const FpParams &p) {
    // static indices only: a dynamically indexed kernel-parameter array would be copied to local memory
    const int lin0 = p.csplit > 1 ? (int)(blockIdx.x / (unsigned)p.csplit) : (int)blockIdx.x;
    if (p.nphase == 0) {   // tile-major, detector blocks fastest (WT_FP_ORDER=tile, kept for A/B measurements)
        FpBlock b;
        b.tile = lin0 / p.ndblk;
        b.dblk = lin0 - b.tile * p.ndblk;
        return b;
    }
    int s0 = 0, s1 = p.phase_start[1];
#pragma unroll
    for (int ph = 1; ph < FP_MAX_PHASES; ++ph)
        if (ph < p.nphase && lin0 >= p.phase_start[ph] * p.ndblk) { s0 = p.phase_start[ph]; s1 = p.phase_start[ph + 1]; }
    const int lin = lin0 - s0 * p.ndblk, nt = s1 - s0;
    FpBlock b;
    b.dblk = lin / nt;
    b.tile = __ldg(p.order + s0 + (lin - b.dblk * nt));
    return b;
}
__device__ __forceinline__ int fp_block_id(const FpParams &p) {   // tile-major id, as (blockIdx.y, blockIdx.x) used to be
    const FpBlock b = fp_block(p);
    return b.tile * p.ndblk + b.dblk;
}

// plain store: sino (n, A, D)
struct FpEpiStore {
    float *sino;
    template <int V>
    __device__ __forceinline__ void operator()(const FpParams &p, int g, int a, int d, bool valid,
                                               const float (&val)[V], float *, int, int) const {
        if (!valid) return;
        const size_t plane = (size_t)p.A * p.D;
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (g * V + i < p.nimg) sino[((size_t)g * V + i) * plane + (size_t)a * p.D + d] = val[i];
    }
};

template <int V>
struct FpSmem {
    typedef typename VecOf<V>::T VecT;
    static constexpr int NS = FpStages<V>::value;
    static_assert(NS >= FP_STAGES, "epilogues count on FP_STAGES stages of scratch");
    VecT seg[NS][FP_LB][FP_W];
    uint64_t full[NS];   // TMA bytes of a stage have landed
    uint64_t empty[NS];  // all consumer warps are done reading a stage
    float wstart[NS];
    float zmin[FP_AT], delta[FP_AT];   // per angle of the tile: left-most ray at line 0 in unwarped coordinates, slope
    // the angles' position warps, entry-major: the producer's 8 lanes (one per angle) read 8 consecutive words
    float wg[16][FP_AT], wi[16][FP_AT];
    int lo, hi;                                    // line range any ray of the CTA needs
};

// float versions of warp_g / warp_ginv over the shared-memory copy of the tile's tables (column `k` = angle slot)
__device__ __forceinline__ float warp_g_f(const float (*wg)[FP_AT], const float (*wi)[FP_AT], int k, float c) {
    const float a = fabsf(c);
    int i = 0;
#pragma unroll
    for (int j = 1; j < WT_NB; ++j)
        if (a >= (float)(2 << j)) i = j;
    const float pos = i ? (float)(2 << i) : 0.0f;
    return copysignf(wg[i][k] + __fdividef(a - pos, wi[i][k]), c);
}
// `si`: the stripe of the previous query (rays move monotonically, a stripe or less per chunk): usually two loads
__device__ __forceinline__ float warp_ginv_f(const float (*wg)[FP_AT], const float (*wi)[FP_AT], int k, float z, int &si) {
    const float a = fabsf(z);
    while (si > 0 && a < wg[si][k]) --si;
    while (si + 1 < WT_NB && a >= wg[si + 1][k]) ++si;
    const float pos = si ? (float)(2 << si) : 0.0f;
    return copysignf(fmaf(a - wg[si][k], wi[si][k], pos), z);
}

// Two consecutive lines of one ray at once -- the float and float2 kernels (one or two images per position), which are
// bound by the instruction issue rate, not by shared-memory wavefronts (11-13 issue slots per line for 1-2 updates).  The
// index arithmetic of the two lines goes through the packed fp32 pipe of sm_100 (FADD2 / FFMA2: one issue slot, each half
// rounded like the scalar operation; x - y is written fma(y, -1, x), whose product is exact), so positions, weights and
// the order of the sum keep their bits: 8 slots per line.  (In the float4 kernel, which is wavefront-bound, the same
// packing measured slower.)  `cr`: c - 1/2 relative to the window start on the two lines; `row`: the first line's window
// minus MAGIC_BITS elements (indexed with the raw bits of the rounded sums: one address instruction per line).
#ifndef WT_FP_PAIRLINES
#define WT_FP_PAIRLINES 1
#endif
template <int V>
__device__ __forceinline__ void fp_line_pair(float2 cr, const typename VecOf<V>::T *row, float (&acc)[V], float &acc_odd) {
    constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23
    constexpr int MAGIC_BITS = 0x4B400000;
    const float2 t = __fadd2_rn(cr, make_float2(MAGIC, MAGIC));
    const float2 u = __fadd2_rn(t, make_float2(-MAGIC, -MAGIC));
    const float2 r = __ffma2_rn(u, make_float2(-1.0f, -1.0f), cr);
    const float2 off = __fadd2_rn(r, make_float2(0.5f, 0.5f));
    const float2 w0 = __ffma2_rn(off, make_float2(-1.0f, -1.0f), make_float2(1.0f, 1.0f));
    check_index(__float_as_int(t.x) - MAGIC_BITS, __float_as_int(t.x) - MAGIC_BITS + 1, FP_W);
    check_index(__float_as_int(t.y) - MAGIC_BITS, __float_as_int(t.y) - MAGIC_BITS + 1, FP_W);
    const typename VecOf<V>::T *ta = row + __float_as_int(t.x), *tb = (row + FP_W) + __float_as_int(t.y);
    float a0[V], a1[V], b0[V], b1[V];
    unpack<V>(ta[0], a0);
    unpack<V>(ta[1], a1);
    unpack<V>(tb[0], b0);
    unpack<V>(tb[1], b1);
    if constexpr (V == 1) {
        // one image: the FMAs of the two lines are packed too, the second line of every pair summing into `acc_odd`
        // (added to the ray sum after the last chunk)
        float2 s2 = make_float2(acc[0], acc_odd);
        s2 = __ffma2_rn(off, make_float2(a1[0], b1[0]), __ffma2_rn(w0, make_float2(a0[0], b0[0]), s2));
        acc[0] = s2.x;
        acc_odd = s2.y;
    } else {
        tap_pair_fma<V>(off.x, a1, w0.x, a0, acc);
        tap_pair_fma<V>(off.y, b1, w0.y, b0, acc);
    }
}

// Epilogue contract: epi.operator()<W>(p, g, a, d, valid, val[W], cta_scratch, pass, npass) handles the W images of
// group g (image index g*W + i) for ray (a, d); it is called by ALL threads of the CTA (it may synchronise).  Under
// MIRROR the V components are V/2 images x {ray (a, d), mirrored ray (a, D-1-d)}: the epilogue then runs twice with
// W = V/2, pass = 0, 1 of npass = 2 (per-CTA partial sums are laid out per pass).
template <int V, class Epi, bool MIRROR>
__global__ void __launch_bounds__(FP_BLOCK) fp_kernel(FpParams p, Epi epi, const __grid_constant__ CUtensorMap tm_vn,
                                                      const __grid_constant__ CUtensorMap tm_vt) {
    typedef typename VecOf<V>::T VecT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FpSmem<V> &sm = *reinterpret_cast<FpSmem<V> *>(smem_raw);

    constexpr int NS = FpStages<V>::value;   // pipeline depth
    constexpr int TAIL_UNROLL = (V <= 2 && WT_FP_PAIRLINES) ? 1 : 4;   // line-by-line loop: all lines (float4), or the odd line left over
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp = 16 detectors x 2 angles; quarter-warp = 4 detectors x 2 angles
    const int slot = (warp >> 1) * 2 + (lane & 1);
    const FpBlock blk = fp_block(p);
    const int d0 = blk.dblk * FP_DT;
    const int d = d0 + (warp & 1) * 16 + (lane >> 3) * 4 + ((lane >> 1) & 3);
    const FpTile tile = p.tiles[blk.tile];
    const int a = tile.first + slot;
    const int g = blockIdx.z;
    const bool producer = warp == FP_THREADS / 32;
    const bool intile = !producer && (slot < tile.count) && (d < p.Dproc);
    const bool valid = intile && (a >= p.a_begin) && (a < p.a_end);   // angle-range calls skip the other rays

    // orientation is uniform over the tile
    const FpAngle an0 = p.ang[tile.first];
    const int nlines = an0.vertical ? p.R : p.C;
    const int npos = an0.vertical ? p.C : p.R;
    const int pitch = fp_pitch(npos);
    const VecT *base = reinterpret_cast<const VecT *>(an0.vertical ? p.vn : p.vt) + (size_t)g * nlines * pitch;

    WT_TRACE_POINT(0, 0, tid == 0);
    WT_TRACE_POINT(0, 1, tid == 0);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], FP_THREADS / 32);
        }
        mbar_fence_init();
        sm.lo = nlines;
        sm.hi = 0;
    }
    if (tid < FP_AT * 8 && (tid >> 3) < tile.count) {    // the angles' position warps: 8 x 16 bytes per angle
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p.warpf + tile.first + (tid >> 3)) + (tid & 7));
        float (*dst)[FP_AT] = (tid & 4) ? sm.wi : sm.wg;          // WarpAngleF = gcum[16], islope[16]
        const int e = 4 * (tid & 3), k = tid >> 3;
        dst[e][k] = v.x; dst[e + 1][k] = v.y; dst[e + 2][k] = v.z; dst[e + 3][k] = v.w;
    }
    __syncthreads();
    if (tid < FP_AT) {
        // the left-most of the tile's rays of this angle at line 0 (detectors d0 .. min(d0+DT, D)-1) in unwarped
        // coordinates (G is increasing: the left-most ray stays the left-most)
        float lo_z = 3.0e38f, dl = 0.0f;
        const int aa = tile.first + tid;
        if (tid < tile.count) {
            const FpAngle an = p.ang[aa];
            const int dl_ = min(d0 + FP_DT, p.Dproc) - 1;
            const double ca = an.c0_base + ((double)d0 + 0.5) * an.c0_step;
            const double cb = an.c0_base + ((double)dl_ + 0.5) * an.c0_step;
            lo_z = warp_g_f(sm.wg, sm.wi, tid, (float)fmin(ca, cb));
            dl = an.delta;
        }
        sm.zmin[tid] = lo_z;
        sm.delta[tid] = dl;
    }

    float c0h = 0.0f, delta = 0.0f, len = 0.0f, c_run = 0.0f;
    int l_lo = 0, l_hi = 0;
    if (intile) {
        const FpAngle an = p.ang[a];
        const float c0 = (float)(an.c0_base + ((double)d + 0.5) * an.c0_step);
        delta = an.delta;
        len = an.len;
        c0h = c0 - 0.5f;
        if (p.astra) {
            // the reference's float32 evaluation of the position at line 0, operation by operation
            // (oracle/astra_linear.c:88-92,119; no contraction)
            const float dh = __fadd_rn((float)d, 0.5f);
            const float Dx = __fadd_rn(an.sx, __fmul_rn(dh, an.ux)), Dy = __fadd_rn(an.sy, __fmul_rn(dh, an.uy));
            if (an.vertical) c_run = __fsub_rn(__fadd_rn(Dx, __fmul_rn(__fsub_rn(p.Ey, Dy), an.ratio)), p.Ex);
            else c_run = -__fsub_rn(__fadd_rn(Dy, __fmul_rn(__fsub_rn(p.Ex, Dx), an.ratio)), p.Ey);
        }
        // lines on which floor(c) can be inside [-1, npos-1], from the warped closed form G(c_l) = G(c_0) + l delta;
        // the estimate may be off by a line on each side (and by the 5e-3 pixel residual of the warp), which the
        // zero padding absorbs (|delta| <= 1, VPAD = 4)
        const float z0 = warp_g_f(sm.wg, sm.wi, slot, c0), zn = warp_g_f(sm.wg, sm.wi, slot, (float)npos);
        if (fabsf(delta) * (float)nlines < 0.5f) {
            const bool hit = (c0 > -2.0f) && (c0 < (float)npos + 1.0f);
            l_lo = 0;
            l_hi = hit ? nlines : 0;
        } else {
            const float inv = 1.0f / delta;
            const float t0 = (-1.05f - z0) * inv, t1 = (zn + 0.05f - z0) * inv;
            // (a ray that starts where the reference's sum does not move has |z0| ~ 1e10: clamp before converting)
            l_lo = (int)fminf(fmaxf(floorf(fminf(t0, t1)), 0.0f), (float)nlines);
            l_hi = (int)fminf(fmaxf(ceilf(fmaxf(t0, t1)) + 1.0f, 0.0f), (float)nlines);
        }
        // chunk boundaries come from ALL rays of the tile, so that a ray's rounding sequence (and result bits)
        // does not depend on which angle range a call asks for
        if (l_hi > l_lo) {
            atomicMin(&sm.lo, l_lo);
            atomicMax(&sm.hi, l_hi);
        }
        if (!valid || l_hi <= l_lo) l_lo = l_hi = 0;    // (a ray that misses the volume has nothing to skip to either)
    }
    __syncthreads();
    WT_TRACE_POINT(0, 2, tid == 0);
    const int Lb = sm.lo, Le = sm.hi;
    const bool tile_in_range = tile.first < p.a_end && tile.first + tile.count > p.a_begin;   // CTA-uniform
    const int nchunk_all = (tile_in_range && Le > Lb) ? (Le - Lb + FP_LB - 1) / FP_LB : 0;
    // Small problems (fewer CTAs than the SMs hold): the CTAs of a cluster share the rays, each walks a range of the line
    // chunks, and rank 0 adds the partial ray sums in rank order before the epilogue (deterministic; the epilogue sees
    // the complete sum).  `ch0`: this rank's first chunk.
    const int cs = p.csplit;
    const int crank = cs > 1 ? (int)(blockIdx.x % (unsigned)cs) : 0;
    int ch0 = 0, nchunk = nchunk_all;
    if (cs > 1) {
        const int per = (nchunk_all + cs - 1) / cs;
        ch0 = min(nchunk_all, crank * per);
        nchunk = min(nchunk_all, ch0 + per) - ch0;
    }

    // producer warp: where the window of chunk `ch` (the ch-th chunk of FP_LB lines, ascending: the reference's
    // summation order) starts -- computed BEFORE the producer waits for the stage to be released
    int stripe = 0;          // producer lanes: stripe of the lane's angle at the last planned chunk
    auto plan = [&](int ch) {
        const int l0 = Lb + (ch0 + ch) * FP_LB;
        const int l1 = min(l0 + FP_LB, nlines) - 1;
        // left-most position any ray of the tile takes on these lines (monotone in l: attained at an end)
        float m = 3.0e38f;
        if (lane < FP_AT) {
            const float zm = sm.zmin[lane], dl = sm.delta[lane];
            if (zm < 1.0e38f) {
                const float za = fmaf((float)l0, dl, zm), zb = fmaf((float)l1, dl, zm);
                if (p.astra) m = fminf(warp_ginv_f(sm.wg, sm.wi, lane, za, stripe), warp_ginv_f(sm.wg, sm.wi, lane, zb, stripe));
                else m = fminf(za, zb);          // the warp is the identity
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        m = __shfl_sync(0xffffffffu, m, 0);
        int ws = (int)floorf(m) - 2;
        ws = max(ws, -VPAD);
        ws = min(ws, pitch - VPAD - FP_W);
        return (ws + VPAD) & ~(4 / V - 1);          // slot index (>= 0), 16-byte aligned
    };
    auto issue = [&](int ch, int ws) {
        const int s = ch % NS;
        const int l0 = Lb + (ch0 + ch) * FP_LB;
        const int nl = min(l0 + FP_LB, nlines) - l0;
        if (lane == 0) {
            fence_proxy_async();                     // consumers' generic-proxy reads of this stage precede the refill
            sm.wstart[s] = (float)(ws - VPAD);       // position of window slot 0
            // (a tiled copy always delivers the whole box: lines past the volume's end arrive as zeros or as the next
            // group's lines and are never read)
            mbar_arrive_expect_tx(&sm.full[s], (uint32_t)((WT_FP_TENSOR_TMA ? FP_LB : nl) * FP_W * sizeof(VecT)));
        }
        __syncwarp();
        if constexpr (WT_FP_TENSOR_TMA) {
            if (lane == 0) {
                const CUtensorMap *tm = an0.vertical ? &tm_vn : &tm_vt;
                const int line = g * nlines + l0;
                // coordinates in elements of the map: floats for V = 1, 2; 8-byte words for V = 4 (fp_tensor_map)
                tma_load_tile_2d(&sm.seg[s][0][0], tm, V == 4 ? ws * 2 : ws * V, line, &sm.full[s]);
            }
        } else {
            if (lane < nl)
                tma_load_1d(&sm.seg[s][lane][0], base + (size_t)(l0 + lane) * pitch + ws, (uint32_t)(FP_W * sizeof(VecT)),
                            &sm.full[s]);
        }
    };

    float acc[V], acc_odd = 0.0f;   // (acc_odd: fp_line_pair<1>)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.0f;
    constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23: adding it rounds to the nearest integer
    constexpr int MAGIC_BITS = 0x4B400000;

    if (producer) {
        // runs ahead of the consumers by up to NS chunks; a stage is refilled as soon as all 8 consumer
        // warps have released it -- consumer warps never wait for each other, only for data
        for (int ch = 0; ch < nchunk; ++ch) {
            const int ws = plan(ch);
            if (ch >= NS) mbar_wait(&sm.empty[ch % NS], (uint32_t)(((ch / NS) - 1) & 1));
            issue(ch, ws);
        }
    } else if (!MIRROR && p.astra) {
        // the reference's running sum: c at line l is c_0 after l float32 additions of delta.  Lines before the ray
        // enters the volume are skipped by doing just the additions.
        if (nchunk > 0) {
            const int lskip = min(l_hi, max(l_lo, Lb + ch0 * FP_LB));      // (a later rank of a cluster starts further down)
#pragma unroll 8
            for (int l = 0; l < lskip; ++l) c_run += delta;
        }
        for (int ch = 0; ch < nchunk; ++ch) {
            const int s = ch % NS;
            mbar_wait(&sm.full[s], (uint32_t)((ch / NS) & 1));
            WT_TRACE_POINT(0, 3, tid == 0 && ch == 0);
            const int l0 = Lb + (ch0 + ch) * FP_LB;
            const int la = max(l0, l_lo), lb = min(l0 + FP_LB, l_hi);
            if (la < lb) {
                // c - 1/2 relative to the window start: window starts are integers, so the difference of two float32
                // values a few positions apart is exact
                const float wsh = sm.wstart[s] + 0.5f;
                const VecT *row = &sm.seg[s][la - l0][0];
                int l = la;
                if constexpr (V <= 2 && WT_FP_PAIRLINES) {
                    const VecT *rowb = row - MAGIC_BITS;
                    // (a separate, fully unrolled path for whole chunks -- no loop control, constant offsets -- measured
                    // slower from 1024^2 on: 3.83 vs 3.65 ms at 2048^2 x 1800, one image)
#pragma unroll 4
                    for (; l + 1 < lb; l += 2) {
                        const float ca = c_run, cb = ca + delta;
                        c_run = cb + delta;
                        fp_line_pair<V>(__fadd2_rn(make_float2(ca, cb), make_float2(-wsh, -wsh)), rowb, acc, acc_odd);
                        rowb += 2 * FP_W;
                    }
                    row = rowb + MAGIC_BITS;
                }
#pragma unroll(TAIL_UNROLL)
                for (; l < lb; ++l) {
                    const float cr = c_run - wsh;
                    const float t = cr + MAGIC;             // round(c - 0.5) = floor(c) (ties: either neighbour, same value)
                    const float off = (cr - (t - MAGIC)) + 0.5f;
                    const int idx = __float_as_int(t) - MAGIC_BITS;
                    check_index(idx, idx + 1, FP_W);
                    float v0[V], v1[V];
                    unpack<V>(row[idx], v0);
                    unpack<V>(row[idx + 1], v1);
                    const float w0 = 1.0f - off;
                    tap_pair_fma<V>(off, v1, w0, v0, acc);
                    c_run += delta;
                    row += FP_W;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);   // this warp is done with stage s
        }
    } else {
        for (int ch = 0; ch < nchunk; ++ch) {
            const int s = ch % NS;
            mbar_wait(&sm.full[s], (uint32_t)((ch / NS) & 1));
            const int l0 = Lb + (ch0 + ch) * FP_LB;
            const int la = max(l0, l_lo), lb = min(l0 + FP_LB, l_hi);
            if (la < lb) {
                // c - 0.5 relative to the window start at line la; then one FADD per line
                float cr = fmaf((float)la, delta, c0h) - sm.wstart[s];
                const VecT *row = &sm.seg[s][la - l0][0];
                int l = la;
                if constexpr (V <= 2 && WT_FP_PAIRLINES) {
                    const VecT *rowb = row - MAGIC_BITS;
#pragma unroll 4
                    for (; l + 1 < lb; l += 2) {
                        const float cb = cr + delta;
                        fp_line_pair<V>(make_float2(cr, cb), rowb, acc, acc_odd);
                        cr = cb + delta;
                        rowb += 2 * FP_W;
                    }
                    row = rowb + MAGIC_BITS;
                }
#pragma unroll(TAIL_UNROLL)
                for (; l < lb; ++l) {
                    const float t = cr + MAGIC;             // round(c - 0.5) = floor(c) (ties: either neighbour, same value)
                    const float off = (cr - (t - MAGIC)) + 0.5f;
                    const int idx = __float_as_int(t) - MAGIC_BITS;
                    check_index(idx, idx + 1, FP_W);
                    float v0[V], v1[V];
                    unpack<V>(row[idx], v0);
                    unpack<V>(row[idx + 1], v1);
                    const float w0 = 1.0f - off;
                    tap_pair_fma<V>(off, v1, w0, v0, acc);
                    cr += delta;
                    row += FP_W;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);   // this warp is done with stage s
        }
    }

    if constexpr (V == 1) acc[0] += acc_odd;
    WT_TRACE_POINT(0, 4, tid == 0);
    if (cs > 1) {
        // partial ray sums of ranks 1 .. cs-1 travel to rank 0's shared memory: part[rank - 1][component][thread]
        float *part = reinterpret_cast<float *>(smem_raw + sizeof(FpSmem<V>));
        if (crank > 0 && !producer) {
#pragma unroll
            for (int i = 0; i < V; ++i) st_cluster_f32(part + ((crank - 1) * V + i) * FP_THREADS + tid, 0, acc[i]);
        }
        cluster_sync_all();
        WT_TRACE_POINT(0, 5, tid == 0);
        if (crank > 0) return;
        if (!producer)
            for (int k = 1; k < cs; ++k)
#pragma unroll
                for (int i = 0; i < V; ++i) acc[i] += part[((k - 1) * V + i) * FP_THREADS + tid];
    }

    // the staging buffers are free now: epilogues may use them as CTA scratch (after their own __syncthreads)
    float *scratch = reinterpret_cast<float *>(&sm.seg[0][0][0]);
    if constexpr (MIRROR) {
        constexpr int W = V / 2;
        float direct[W], mirrored[W];
#pragma unroll
        for (int i = 0; i < W; ++i) { direct[i] = acc[i] * len; mirrored[i] = acc[W + i] * len; }
        const int dm = p.D - 1 - d;
        epi.template operator()<W>(p, g, a, d, valid, direct, scratch, 0, 2);
        epi.template operator()<W>(p, g, a, dm, valid && dm != d, mirrored, scratch, 1, 2);
    } else {
        float val[V];
#pragma unroll
        for (int i = 0; i < V; ++i) val[i] = acc[i] * len;
        epi.template operator()<V>(p, g, a, d, valid, val, scratch, 0, 1);
    }
    WT_TRACE_POINT(0, 6, tid == 0);
    WT_TRACE_POINT(0, 7, tid == 0);
}

// Tensor map of one packed copy (G groups x nlines lines x pitch positions x V floats) as a rank-2 tensor (position, line)
// with a box of FP_LB lines x FP_W positions.  TMA boxes end at 256 elements per dimension: V = 1, 2 count in floats
// (96 / 192 per line), V = 4 in 8-byte words (192 per line; a rank-3 map with the four components as its innermost
// 16-byte dimension works too but moves the box 16 bytes at a time: 1.5x slower than the one-line copies).
inline cudaError_t fp_tensor_map(CUtensorMap *tm, const float *packed, int V, int pitch, size_t total_lines) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = (EncodeFn)fn;
    }
    const int per_pos = V == 4 ? 2 : V;          // map elements per position
    cuuint64_t dims[2] = {(cuuint64_t)pitch * per_pos, (cuuint64_t)total_lines};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * V * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)(FP_W * per_pos), (cuuint32_t)FP_LB}, estr[2] = {1, 1};
    const CUresult r = encode(tm, V == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)packed, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Cluster split of the line chunks (fp_kernel): two CTAs per (tile, detector block) when there are at most about two
// waves of them (4 CTAs per SM) and every rank keeps at least four chunks -- measured on the C-side SIRT loop: 256^2
// 52 -> 39 us, 384^2 95 -> 88 us, 512^2 161 -> 156 us per iteration; slower from 768^2 on (tools/split_probe*.sh).  It
// depends on the geometry alone -- never on the number of images, the angle range or the epilogue -- so a ray's bits do
// not depend on the call it travels in.
constexpr int FP_MAX_SPLIT = 8;
inline int fp_cluster_split(int blocks, int nlines_max) {
    static const int off = getenv("WT_NO_SPLIT") ? 1 : 0;       // A/B measurements
    if (off) return 1;
    static const int forced = getenv("WT_FP_SPLIT") ? atoi(getenv("WT_FP_SPLIT")) : 0;
    if (forced >= 1 && forced <= FP_MAX_SPLIT) return forced;
    const int nchunk = (nlines_max + FP_LB - 1) / FP_LB;
    return (blocks <= 1200 && nchunk >= 8) ? 2 : 1;
}

template <int V, class Epi, bool MIRROR = false>
inline cudaError_t launch_fp(const FpParams &p, const Epi &epi, int groups, cudaStream_t st) {
    FpParams q = p;
    q.ndblk = (p.Dproc + FP_DT - 1) / FP_DT;
    constexpr size_t part_bytes = (size_t)V * FP_THREADS * sizeof(float);
    const int cs = fp_cluster_split(q.ndblk * p.ntiles, max(p.R, p.C));
    q.csplit = cs;
    const size_t smem = sizeof(FpSmem<V>) + (cs - 1) * part_bytes;
    static bool configured[64] = {};   // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fp_kernel<V, Epi, MIRROR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(FpSmem<V>) + (FP_MAX_SPLIT - 1) * part_bytes));
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    // tensor maps of the two packed copies (host-side encoding of a few hundred bytes; they travel as kernel parameters)
    alignas(64) CUtensorMap tm_vn, tm_vt;
    memset(&tm_vn, 0, sizeof(tm_vn));
    memset(&tm_vt, 0, sizeof(tm_vt));
    if (WT_FP_TENSOR_TMA) {
        cudaError_t e = fp_tensor_map(&tm_vn, p.vn, V, fp_pitch(p.C), (size_t)groups * p.R);
        if (e == cudaSuccess) e = fp_tensor_map(&tm_vt, p.vt, V, fp_pitch(p.R), (size_t)groups * p.C);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(q.ndblk * p.ntiles * cs, 1, groups);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(FP_BLOCK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    LaunchAttrs la;
    if (cs > 1) la.cluster(cs, 1, 1);
    cfg.attrs = la.at;
    cfg.numAttrs = la.n;
    return cudaLaunchKernelEx(&cfg, fp_kernel<V, Epi, MIRROR>, q, epi, tm_vn, tm_vt);
}

}  // namespace wt
