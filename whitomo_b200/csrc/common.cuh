// Shared device helpers and the geometry handle of the whitomo_b200 engine (sm_100a only).
#pragma once
#include <cuda.h>
#include <cstdlib>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <string>

#include "../../include/whitomo_b200.h"

namespace wt {

// ---- layout constants -------------------------------------------------------------------------
// Packed volumes carry VPAD zero positions on both sides of every line so that the two Joseph taps of
// any in-range (and up to 3 positions out-of-range) ray position are always readable: no bounds
// predicates in the projector's inner loop.
constexpr int VPAD = 4;

// Backprojector tile: BP_TW x TH pixels per CTA (TH = 32, or 16 for small volumes, see bp_tile_height), BP_AB angles per
// TMA stage, BP_STAGES stages.
constexpr int BP_TW = 64;
#ifndef WT_BP_STAGES
#define WT_BP_STAGES 3
#endif
#ifndef WT_BP_AB
#define WT_BP_AB 16
#endif
constexpr int BP_AB = WT_BP_AB;
constexpr int BP_STAGES = WT_BP_STAGES;
// staged segment width per tile height (>= tile diagonal + 8, multiple of 4): 8 consumer warps x TH/8 rows per thread
template <int TH> struct BpTile {
    static_assert(TH == 16 || TH == 32, "tile heights: 16 or 32 rows");
    static constexpr int RPT = TH / 8;                 // rows per thread
    static constexpr int SEGW = TH >= 32 ? 80 : 72;
};
constexpr int BP_SEGW_MAX = 80;
constexpr size_t WT_BP_PLAN_TABLE_MAX = 8u << 20;   // bytes of precomputed K2 plans a geometry may own (bp.cuh: BpPlanRec)
constexpr int BP_THREADS = 256;            // consumer threads
constexpr int BP_BLOCK = BP_THREADS + 32;  // + one producer warp that only drives the TMA pipeline

// ---- per-angle constants ------------------------------------------------------------------------
// FP: position along the interpolation axis for line l and detector d is
//     c(l, d) = c0_base + (d + 0.5) * c0_step + l * delta        (ASTRA linear projector)
// with the float32 vector-geometry fields ASTRA would hold, combined in double once per ray.
// The float32 fields themselves (sx .. ratio) let the kernel repeat ASTRA's own float32 evaluation of the position
// at line 0 operation by operation (WT_POS_ASTRA, see fp.cuh).
struct FpAngle {
    double c0_base;
    double c0_step;
    float delta;    // tan(theta) (vertical) or cot(theta) (horizontal), float32 like the reference
    float len;      // 1/|cos| or 1/|sin|
    int vertical;   // |ray_x| < |ray_y| evaluated on float32 values
    float sx, sy;   // detector start  (float32 fields of ASTRA's vector geometry)
    float ux, uy;   // detector step
    float ratio;    // ray_x / ray_y (vertical) or ray_y / ray_x (horizontal)
};

// BP (gather form of the exact transpose, SURVEY A.3, derived from the same FpAngle so that
// BP == FP^T including the float32 rounding of the constants):
//     d*(row, col) = d00 + col * dcol + row * drow;  taps floor(d*), floor(d*)+1 with weights
//     max(0, A0 - B f) and max(0, A1 + B f),  f = frac(d*),  A0 = len, B = len*|c0_step|, A1 = A0 - B.
struct BpAngle {
    double d00;
    double dcol;
    double drow;
    float a0, a1, b;
    float pad_;
};

// ---- the reference's float32 running sum as a position warp (WT_POS_ASTRA) ----------------------------------------
// ASTRA advances a ray's position with a float32 running sum, c += delta once per line (SURVEY A.2;
// oracle/astra_linear.c:93,120).  While |c| stays inside one binade [2^e, 2^(e+1)) every addition rounds to that
// binade's grid u = 2^(e-23), so the ray does not advance by delta but by step_e = rint(delta / u) * u: its speed
// depends on WHERE it is, not on which ray it is.  Integrating dc/dl = step(|c|) gives the "unwarped" coordinate
//     G(c) = int_0^c delta / step(|c'|) dc'      (odd, piecewise linear, slope 1 for |c| < 4, kinks at +-2^e)
// in which every ray moves at the nominal speed again:  G(c_l) = G(c_0) + l * delta.  The reference's position is
// therefore  c_l(d) = Ginv(G(c_0(d)) + l * delta)  up to the O(u) discretisation at the <= 12 binade crossings of a
// ray (measured: 1.5e-4 pixel rms at N = 2048 and 3e-4 at N = 8192, where the plain closed form c_0 + l * delta is
// 3e-2 resp. 0.6 pixel rms -- up to 2.5 pixel -- away from what the reference computes).  K1 uses the table for its
// staging windows and line ranges (its positions themselves repeat the float32 sum literally); K2 inverts the map:
// the detector coordinate of pixel (line l, position k) is  d* = (Ginv(G(k) - l * delta) - c0_base) / c0_step - 1/2.
// Under WT_POS_EXACT every slope is 1 and G is the identity.
constexpr int WT_NB = 15;   // stripes of |position|: [0, 4), [4, 8), ..., [2^15, inf)
constexpr double WT_STUCK_SLOPE = 1048576.0;   // slope of G where the reference's ray does not move at all (|delta| < u/2)
struct WarpAngle {
    double gcum[WT_NB];     // G at the lower edge of stripe i (gcum[0] = 0)
    double slope[WT_NB];    // dG/dc inside stripe i = delta / step_i
    double islope[WT_NB];   // 1 / slope
};
// float copies for the kernels' searches: the fine table (K1's staging windows and line ranges) ...
struct WarpAngleF {
    float gcum[16];
    float islope[16];     // entries [WT_NB, 16) are padding
};
// ... and K2's view: the stripes below 64 merged into one with the secant slope G(64) / 64 (used for G and Ginv alike, so
// the two stay inverse to each other): merged stripe 0 = [0, 64), m >= 1 = stripe m + 4.  cum[m] = G at the lower edge,
// bend[m] = |1/slope_m - 1/slope_(m-1)| (how much d* bends at that edge); a0, b: the tap-weight constants of the angle.
constexpr int WT_NM = WT_NB - 4;
struct BpWarpF {
    float cum[WT_NM];
    float a0;
    float bend[WT_NM];
    float b;
    float slope[WT_NM];
    float inv_step;         // 1 / c0_step
    float isl[WT_NM];       // 1 / slope
    float delta;
};
// the same merged view in double, with everything else K2's producer needs of an angle (no division on the device)
struct BpWarpD {
    double cum[WT_NM];
    double slope[WT_NM];
    double isl[WT_NM];      // 1 / slope
    double c0_base, inv_step, delta;
    int vertical, pad_;
};
__host__ __device__ inline double stripe_pos(int i) { return i == 0 ? 0.0 : (double)(2 << i); }
__host__ __device__ inline int stripe_of(double a) {   // a >= 0
    int i = 0;
    for (int j = 1; j < WT_NB; ++j)
        if (a >= (double)(2 << j)) i = j;
    return i;
}
__host__ __device__ inline double warp_g(const WarpAngle &w, double c) {
    const double a = fabs(c);
    const int i = stripe_of(a);
    const double g = w.gcum[i] + (a - stripe_pos(i)) * w.slope[i];
    return c < 0.0 ? -g : g;
}
__host__ __device__ inline double warp_ginv(const WarpAngle &w, double z) {
    const double a = fabs(z);
    int i = 0;
    for (int j = 1; j < WT_NB; ++j)
        if (a >= w.gcum[j]) i = j;
    const double c = stripe_pos(i) + (a - w.gcum[i]) * w.islope[i];
    return z < 0.0 ? -c : c;
}

}  // namespace wt

struct wt_geom {
    int rows, cols, ndet, nangles;
    float det_width;
    int sino_padl;   // zero detectors in front of every packed sinogram row
    int sino_pitch;  // packed sinogram row pitch in detector positions (multiple of 4)
    wt::FpAngle *d_fp;
    wt::BpAngle *d_bp;
    wt::WarpAngle *d_warp;   // per angle: the reference's float32 running sum as a position warp (identity under WT_POS_EXACT)
    wt::WarpAngleF *d_warpf; // float copy for K1
    wt::BpWarpF *d_bwarp;    // merged float view for K2
    wt::BpWarpD *d_bwarpd;   // ... and its double version
    void *d_bplans;          // wt::BpPlanRec[tiles][nangles]: K2's per-(tile, angle) plans, small geometries only (else null)
    int positions;           // WT_POS_ASTRA | WT_POS_EXACT
    void *d_tiles;   // wt::FpTile[ntiles]: angle tiles of the forward projector
    int ntiles;
    int *d_order;    // launch order of the tiles, phase by phase (fp.cuh: fp_block)
    int nphase;
    int phase_start[5];
    float *h_angles;
    unsigned char *h_vertical;
    // optional device timestamps of the fused step (WT_STEP_TIMED): ring of (before K2, after K2, after K1) events
    cudaEvent_t *step_ev;
    long long step_count;   // per angle: 1 if the projector marches over rows (reads VN), 0 if over columns (VT)
};

struct wt_spectrum {
    int K, E;
    float *d_tab;  // [E] s_e*w_e, then K rows [E] of pix*ea[k,e]*log2(e)  (the kernels use exp2)
    int cslot;     // slot of the __constant__ table pool on `device`, or -1 (epilogues.cuh: c_spectrum_tab)
    int device;
};

namespace wt {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define WT_CUDA(call)                                                 \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return wt::cuda_fail(e__, #call);     \
    } while (0)

// ---- small vector helpers (V images interleaved per position) -----------------------------------
template <int V> struct VecOf;
template <> struct VecOf<1> { typedef float T; };
template <> struct VecOf<2> { typedef float2 T; };
template <> struct VecOf<4> { typedef float4 T; };

template <int V> __device__ __forceinline__ void unpack(const typename VecOf<V>::T &t, float (&v)[V]);
template <> __device__ __forceinline__ void unpack<1>(const float &t, float (&v)[1]) { v[0] = t; }
template <> __device__ __forceinline__ void unpack<2>(const float2 &t, float (&v)[2]) { v[0] = t.x; v[1] = t.y; }
template <> __device__ __forceinline__ void unpack<4>(const float4 &t, float (&v)[4]) {
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

template <int V> __device__ __forceinline__ typename VecOf<V>::T pack(const float (&v)[V]);
template <> __device__ __forceinline__ float pack<1>(const float (&v)[1]) { return v[0]; }
template <> __device__ __forceinline__ float2 pack<2>(const float (&v)[2]) { return make_float2(v[0], v[1]); }
template <> __device__ __forceinline__ float4 pack<4>(const float (&v)[4]) {
    return make_float4(v[0], v[1], v[2], v[3]);
}

// a[i] = fma(w0, v0[i], fma(w1, v1[i], a[i])) for the V interleaved images of one tap pair.  With four images the
// two halves go through the packed FFMA2 of sm_100 (two fp32 FMAs per issue slot, each rounded like a scalar fma, so
// the result bits are those of the scalar form): K1 and K2 spend half of their issue slots on these FMAs.
#ifndef WT_FFMA2
#define WT_FFMA2 1
#endif
template <int V>
__device__ __forceinline__ void tap_pair_fma(float w0, const float (&v0)[V], float w1, const float (&v1)[V], float (&a)[V]) {
    if constexpr (V == 4 && WT_FFMA2) {
        const float2 W0 = make_float2(w0, w0), W1 = make_float2(w1, w1);
        float2 lo = make_float2(a[0], a[1]), hi = make_float2(a[2], a[3]);
        lo = __ffma2_rn(W0, make_float2(v0[0], v0[1]), __ffma2_rn(W1, make_float2(v1[0], v1[1]), lo));
        hi = __ffma2_rn(W0, make_float2(v0[2], v0[3]), __ffma2_rn(W1, make_float2(v1[2], v1[3]), hi));
        a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
    } else if constexpr (V == 2 && WT_FFMA2) {
        float2 lo = make_float2(a[0], a[1]);
        lo = __ffma2_rn(make_float2(w0, w0), make_float2(v0[0], v0[1]), __ffma2_rn(make_float2(w1, w1), make_float2(v1[0], v1[1]), lo));
        a[0] = lo.x; a[1] = lo.y;
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = fmaf(w0, v0[i], fmaf(w1, v1[i], a[i]));
    }
}

// a[i] = fma(w0, v0[i], fma(w1, v1[i], fma(w2, v2[i], a[i])))
template <int V>
__device__ __forceinline__ void tap_triple_fma(float w0, const float (&v0)[V], float w1, const float (&v1)[V], float w2,
                                               const float (&v2)[V], float (&a)[V]) {
    if constexpr (V == 4 && WT_FFMA2) {
        const float2 W0 = make_float2(w0, w0), W1 = make_float2(w1, w1), W2 = make_float2(w2, w2);
        float2 lo = make_float2(a[0], a[1]), hi = make_float2(a[2], a[3]);
        lo = __ffma2_rn(W0, make_float2(v0[0], v0[1]), __ffma2_rn(W1, make_float2(v1[0], v1[1]), __ffma2_rn(W2, make_float2(v2[0], v2[1]), lo)));
        hi = __ffma2_rn(W0, make_float2(v0[2], v0[3]), __ffma2_rn(W1, make_float2(v1[2], v1[3]), __ffma2_rn(W2, make_float2(v2[2], v2[3]), hi)));
        a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = fmaf(w0, v0[i], fmaf(w1, v1[i], fmaf(w2, v2[i], a[i])));
    }
}

// ---- optional bounds checks of the shared-memory gathers (-DWT_BOUNDS_CHECK=1, see wt_debug_oob_count) -------------
#ifndef WT_BOUNDS_CHECK
#define WT_BOUNDS_CHECK 0
#endif
__device__ unsigned long long g_oob_count;
__device__ __forceinline__ void check_index(int lo, int hi, int limit) {   // the gather reads slots [lo, hi] of [0, limit)
#if WT_BOUNDS_CHECK
    if (lo < 0 || hi >= limit) atomicAdd(&g_oob_count, 1ull);
#endif
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + TMA (1D bulk copy) ----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// A warp whose data has not arrived yields its issue slots for a moment instead of spinning: on small problems the
// consumers otherwise starve the one warp that could feed them (the producer shares their schedulers).
#ifndef WT_WAIT_SLEEP_NS
#define WT_WAIT_SLEEP_NS 32
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WT_DONE_%=;\n"
        "WT_WAIT_%=:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WT_WAIT_%=;\n"
        "WT_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "n"(WT_WAIT_SLEEP_NS)
        : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar` in bytes.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// global -> shared tiled copies (one box of a tensor map); coordinates innermost first
__device__ __forceinline__ void tma_load_tile_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_tile_3d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// ---- development aid (-DWT_TRACE=1 builds only; tools/trace_small.py): per-CTA time stamps of K1 / K2 ----
#ifndef WT_TRACE
#define WT_TRACE 0
#endif
#if WT_TRACE
constexpr int WT_TRACE_CTAS = 2048, WT_TRACE_POINTS = 8;
__device__ unsigned long long g_wt_trace[2][WT_TRACE_CTAS][WT_TRACE_POINTS];   // [kernel][linear CTA][point]: point 0 = globaltimer (ns), others clock64
__device__ __forceinline__ void wt_trace_point(int kernel, int point) {
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (cta >= (unsigned)WT_TRACE_CTAS) return;
    unsigned long long t;
    if (point == 0 || point == WT_TRACE_POINTS - 1) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    else t = clock64();
    g_wt_trace[kernel][cta][point] = t;
}
#define WT_TRACE_POINT(kernel, point, cond) do { if (cond) wt::wt_trace_point(kernel, point); } while (0)
#else
#define WT_TRACE_POINT(kernel, point, cond) do { } while (0)
#endif

// launch attributes of K1 / K2: the optional cluster dimension.  (Programmatic dependent launch -- letting a kernel's
// set-up overlap the previous kernel's tail -- was built and measured in round 2: the CTAs did start early and wait, but
// neither the 256^2 SIRT loop (25.15k vs 25.16k iterations/s) nor the C4 step changed; removed.)
struct LaunchAttrs {
    cudaLaunchAttribute at[1];
    int n = 0;
    void cluster(int x, int y, int z) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = x; at[n].val.clusterDim.y = y; at[n].val.clusterDim.z = z;
        ++n;
    }
};

// ---- thread-block clusters: small problems split a tile's angles (K2) / lines (K1) over the CTAs of a cluster ----
// All threads of all CTAs of the cluster that have not exited: release their writes (remote ones included), wait, acquire.
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();      // (.aligned: the warp must arrive converged)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the shared memory of CTA `rank` of the cluster, at the address `local` has in this CTA
__device__ __forceinline__ void st_cluster_f32(const float *local, uint32_t rank, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace wt
