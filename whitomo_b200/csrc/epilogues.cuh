// Fused epilogues of K1/K2 and the streaming kernels around them (K3 and small reductions).
#pragma once
#include "bp.cuh"
#include "common.cuh"
#include "fp.cuh"

namespace wt {

// ---- SIRT (ASTRA CSirtAlgorithm): d = (p - W x) / Rsum, x += (W^T d) / Csum, guarded -----
struct FpEpiSirtResidual {
    const float *meas;    // (n, A, D)
    const float *raysum;  // (A, D)
    float *out;           // (n, A, D), or null when the residual goes straight into K2's packed layout:
    float *sp;            // (G, A, pitch, V) with `padl` zero detectors in front (bp.cuh: pack_sino_kernel), or null
    int padl, pitch;
    template <int V>
    __device__ __forceinline__ void operator()(const FpParams &p, int g, int a, int d, bool valid,
                                               const float (&val)[V], float *, int pass, int) const {
        if (!valid) return;
        const size_t plane = (size_t)p.A * p.D, r = (size_t)a * p.D + d;
        const float rs = __ldg(raysum + r);
        float res[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            res[i] = 0.0f;
            if (g * V + i >= p.nimg) continue;       // padding component of the last group
            const size_t o = ((size_t)g * V + i) * plane + r;
            const float diff = __ldg(meas + o) - val[i];
            res[i] = rs > 0.000001f ? diff / rs : 0.0f;
            if (out) out[o] = res[i];
        }
        if (sp && pass == 0)    // (packed output is never combined with mirror pairing)
            reinterpret_cast<typename VecOf<V>::T *>(sp)[((size_t)g * p.A + a) * pitch + padl + d] = pack<V>(res);
    }
};

struct BpEpiSirtUpdate {
    static constexpr bool kHasSums = false;
    static constexpr int kMult = 1;
    float *vol;           // (n, R, C), updated in place
    const float *pixsum;  // (R, C)
    float *vn, *vt;       // K1's packed copies of the volume, refreshed in place (fp.cuh: pack_vol_kernel), or null
    template <int V>
    __device__ __forceinline__ void operator()(const BpParams &p, int g, int row, int col,
                                               const float (&val)[V]) const {
        const size_t img = (size_t)p.R * p.C, px = (size_t)row * p.C + col;
        const float cs = __ldg(pixsum + px);
        float nv[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            nv[i] = 0.0f;
            if (g * V + i >= p.nimg) continue;       // padding component of the last group
            const size_t o = ((size_t)g * V + i) * img + px;
            nv[i] = vol[o] + (cs > 0.000001f ? val[i] / cs : 0.0f);
            vol[o] = nv[i];
        }
        if (vn) {
            typedef typename VecOf<V>::T VecT;
            reinterpret_cast<VecT *>(vn)[((size_t)g * p.R + row) * fp_pitch(p.C) + VPAD + col] = pack<V>(nv);
            reinterpret_cast<VecT *>(vt)[((size_t)g * p.C + col) * fp_pitch(p.R) + VPAD + row] = pack<V>(nv);
        }
    }
};

__global__ void fill_kernel(float *p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- white-beam model fused into the projector (src/white_projection.py:86-98,116-139) -----------------
#ifndef WT_CTAB
#define WT_CTAB 384   // floats per constant-memory slot (E <= 128 at K = 2)
#endif
constexpr int WT_CSLOTS = 16;   // live spectra with a constant-memory slot (per device)
__constant__ float c_spectrum_tab[WT_CSLOTS * WT_CTAB];

// exp2 for the bin loop: one MUFU.EX2, results below the normal range flushed to zero (such a bin contributes
// nothing to a float32 sum of 64 terms of order one)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// A group of V interleaved images holds V/K slices x K elements (image index = slice*K + k).
template <int K>
struct FpEpiWhite {
    const float *tab;   // [E] s_e*w_e, then K x [E] pix*ea[k][e]*log2(e)
    int cslot;          // slot of c_spectrum_tab holding the same table, or -1
    int E;
    const float *meas;  // (S, A, D) or null
    float *inten;       // (S, A, D) or null
    float *qmu;         // (S, K, A, D) or null
    float *partial;     // (S, nblocks) per-CTA sums of q^2, or null
    float *sp;          // q*mu in K2's packed layout (G, A, pitch, V), `padl` zero detectors in front, or null
    int padl, pitch;
    template <int V>
    __device__ __forceinline__ void operator()(const FpParams &p, int g, int a, int d, bool valid,
                                               const float (&val)[V], float *scratch, int pass, int npass) const {
        static_assert(V % K == 0, "group must hold whole slices");
        constexpr int SPG = V / K;
        __shared__ float red[SPG][FP_BLOCK / 32];
        const size_t plane = (size_t)p.A * p.D, r = (size_t)a * p.D + d;
        const int nblocks = p.ndblk * p.ntiles * npass;
        const int bid = pass * p.ndblk * p.ntiles + fp_block_id(p);
        const int tid = threadIdx.x;
        // Spectrum tables.  A spectrum of up to WT_CTAB floats owns a slot of __constant__ memory (filled once, at
        // wt_spectrum_create): the bin loop reads it through the constant cache (LDC with a warp-uniform index) and
        // costs no shared-memory wavefronts -- K + 1 broadcast LDS per bin were 3.5 % of this LSU-bound kernel's
        // wavefronts (a float4 row per bin was worse: a broadcast LDS.128 still takes four).  Larger tables, or more
        // live spectra than slots, use the idle staging buffers, or stay in global memory.
        if (partial && pass > 0) __syncthreads();   // the previous pass may still be summing `red` (CTA-uniform condition)
        const int ntab = (K + 1) * E;
        const bool in_const = cslot >= 0;
        const bool in_smem = !in_const && ntab * (int)sizeof(float) <= FP_STAGES * FP_LB * FP_W * (int)sizeof(float);
        if (!in_const) {
            __syncthreads();
            if (in_smem)
                for (int i = tid; i < ntab; i += FP_BLOCK) scratch[i] = __ldg(tab + i);
            __syncthreads();
        }
        const float *tb = in_smem ? scratch : tab;
        // The bin loop runs once over the tables for all slices of the group (the table reads are shared), and once
        // per table source so that the constant-memory version compiles to LDC (a pointer chosen at run time would
        // turn both into generic loads).  Padding slices compute on zeros and are not stored.
        float I[SPG], mu[SPG][K];
#pragma unroll
        for (int sl = 0; sl < SPG; ++sl) {
            I[sl] = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) mu[sl][k] = 0.0f;
        }
        auto bins = [&](auto table) {
            for (int e = 0; e < E; ++e) {
                const float sw = table(e);
                float ab[K];
#pragma unroll
                for (int k = 0; k < K; ++k) ab[k] = table((k + 1) * E + e);
#pragma unroll
                for (int sl = 0; sl < SPG; ++sl) {
                    float arg = 0.0f;
#pragma unroll
                    for (int k = 0; k < K; ++k) arg = fmaf(ab[k], val[sl * K + k], arg);
                    const float t = sw * ex2_ftz(-arg);       // tables are pre-scaled by log2(e): one MUFU.EX2
                    I[sl] += t;
#pragma unroll
                    for (int k = 0; k < K; ++k) mu[sl][k] = fmaf(t, ab[k], mu[sl][k]);
                }
            }
        };
        if (valid) {
            const int cbase = in_const ? cslot * WT_CTAB : 0;
            if (in_const) bins([&](int i) { return c_spectrum_tab[cbase + i]; });
            else bins([&](int i) { return tb[i]; });
        }
        float packed[V];
#pragma unroll
        for (int i = 0; i < V; ++i) packed[i] = 0.0f;
#pragma unroll
        for (int sl = 0; sl < SPG; ++sl) {
            const size_t slice = (size_t)g * SPG + sl;
            const bool real = (g * SPG + sl) * K < p.nimg;      // else: padding slice of the last group
            float q2 = 0.0f;
            if (valid && real) {
                const float q = (meas ? __ldg(meas + slice * plane + r) : 0.0f) - I[sl];
                if (inten) inten[slice * plane + r] = I[sl];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    packed[sl * K + k] = q * (-0.6931471805599453f * mu[sl][k]);
                    if (qmu) qmu[(slice * K + k) * plane + r] = packed[sl * K + k];
                }
                q2 = q * q;
            }
            if (partial) {
                q2 = warp_sum(q2);
                if ((tid & 31) == 0) red[sl][tid >> 5] = q2;
            }
        }
        if (sp && valid && npass == 1)   // (packed output is never combined with mirror pairing)
            reinterpret_cast<typename VecOf<V>::T *>(sp)[((size_t)g * p.A + a) * pitch + padl + d] = pack<V>(packed);
        if (partial) {
            __syncthreads();
            if (tid < SPG && (g * SPG + tid) * K < p.nimg) {
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < FP_BLOCK / 32; ++w) s += red[tid][w];
                partial[((size_t)g * SPG + tid) * nblocks + bid] = s;
            }
        }
    }
};

// ---- mono MAR residual fused into the projector (teeth_recon/experiments.py:712-722, teeth_recon/barrier.py:50-69)
struct FpEpiMar {
    const float *b;              // (n, A, D)
    const unsigned char *mask;   // (n, A, D)
    const float *inv_ngood;      // (n)
    float bt, b_hb;
    float *u;                    // (n, A, D) or null
    float *proj;                 // (n, A, D) or null
    float *partial;              // (3, n, nblocks) or null
    int nimg;                    // n of this launch (stride of the partial planes)
    template <int V>
    __device__ __forceinline__ void operator()(const FpParams &p, int g, int a, int d, bool valid,
                                               const float (&val)[V], float *, int pass, int npass) const {
        __shared__ float red[3][V][FP_BLOCK / 32];
        const size_t plane = (size_t)p.A * p.D, r = (size_t)a * p.D + d;
        const int nblocks = p.ndblk * p.ntiles * npass;
        const int bid = pass * p.ndblk * p.ntiles + fp_block_id(p);
        const int tid = threadIdx.x;
        if (partial) __syncthreads();   // a previous pass may still be reading `red`
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const size_t im = (size_t)g * V + i;
            float cost = 0.0f, phi = 0.0f, bad = 0.0f;
            if (valid && g * V + i < p.nimg) {
                const size_t o = im * plane + r;
                const float P = val[i];
                float uu;
                if (mask[o]) {
                    const float fx = P - b_hb;                      // = -(b_hb - P)
                    uu = bt != 0.0f ? -bt / fx : 0.0f;
                    phi = fx > 0.0f ? -logf(fx) : INFINITY;
                    bad = P < b_hb ? 1.0f : 0.0f;
                } else {
                    const float dd = (P - __ldg(b + o)) * __ldg(inv_ngood + im);
                    uu = 2.0f * dd;
                    cost = dd * dd;
                }
                if (u) u[o] = uu;
                if (proj) proj[o] = P;
            }
            if (partial) {
                cost = warp_sum(cost); phi = warp_sum(phi); bad = warp_sum(bad);
                if ((tid & 31) == 0) { red[0][i][tid >> 5] = cost; red[1][i][tid >> 5] = phi; red[2][i][tid >> 5] = bad; }
            }
        }
        if (partial) {
            __syncthreads();
            if (tid < 3 * V && g * V + tid % V < p.nimg) {
                const int q = tid / V, i = tid % V;
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < FP_BLOCK / 32; ++w) s += red[q][i][w];
                partial[((size_t)q * nimg + (size_t)g * V + i) * nblocks + bid] = s;
            }
        }
    }
};

// generic-K path: P (S, K, A, D) already projected; one thread per ray
__global__ void white_epilogue_kernel(const float *__restrict__ P, const float *__restrict__ tab, int K, int E,
                                      const float *__restrict__ meas, float *inten, float *qmu, float *partial,
                                      size_t plane, size_t r_begin, size_t r_end) {
    const size_t slice = blockIdx.y;
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float q2 = 0.0f;
    if (r >= r_begin && r < r_end) {
        float I = 0.0f;
        for (int e = 0; e < E; ++e) {
            float arg = 0.0f;
            for (int k = 0; k < K; ++k) arg = fmaf(__ldg(tab + (k + 1) * E + e), __ldg(P + (slice * K + k) * plane + r), arg);
            I += __ldg(tab + e) * exp2f(-arg);
        }
        const float q = (meas ? __ldg(meas + slice * plane + r) : 0.0f) - I;
        if (inten) inten[slice * plane + r] = I;
        if (qmu) {
            for (int k = 0; k < K; ++k) {
                float mu = 0.0f;
                for (int e = 0; e < E; ++e) {
                    float arg = 0.0f;
                    for (int kk = 0; kk < K; ++kk)
                        arg = fmaf(__ldg(tab + (kk + 1) * E + e), __ldg(P + (slice * K + kk) * plane + r), arg);
                    mu = fmaf(__ldg(tab + e) * exp2f(-arg), __ldg(tab + (k + 1) * E + e), mu);
                }
                qmu[(slice * K + k) * plane + r] = q * (-0.6931471805599453f * mu);
            }
        }
        q2 = q * q;
    }
    if (partial) {
        __shared__ float red[8];
        q2 = warp_sum(q2);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q2;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.0f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            partial[slice * gridDim.x + blockIdx.x] = s;
        }
    }
}

// out[s] = scale * sum_b partial[s][b]   (fixed order, float64) -- one warp per slice
__global__ void reduce_partials_kernel(const float *__restrict__ partial, int nb, double scale, double *out) {
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nb; b += 32) acc += (double)partial[(size_t)s * nb + b];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[s] = scale * acc;
}
__global__ void reduce_partials_strided_kernel(const float *__restrict__ partial, int nb, double *out, int stride) {
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nb; b += 32) acc += (double)partial[(size_t)s * nb + b];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[(size_t)s * stride] = acc;
}
__global__ void reduce_partials_f64_kernel(const double *__restrict__ partial, int nb, double scale, double *out) {
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nb; b += 32) acc += partial[(size_t)s * nb + b];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[s] = scale * acc;
}

// ---- K3: fused box-barrier gradient step (src/barrier_method.py:164-206, gogo constraints) -------------
// One thread per pixel handles all K elements of one slice (K <= 4).  HBM-bound: reads x, g; writes x.
constexpr int UPD_BLOCKS_PER_SLICE = 296;  // 2 CTAs per SM worth of partials per slice

struct UpdFlags {
    bool act, clip, use_reg, use_bar, lower_only, want_pen, use_triv, all_elems, fast_log;
    float bt;
};

// one pixel, all K elements: returns the penalty contribution of the new point
template <int K>
__device__ __forceinline__ float upd_pixel(const wt_update_params &prm, const UpdFlags &f, const float (&xv)[K],
                                           const float (&gv)[K], float (&xn)[K]) {
    float sgn = 0.0f;
    if (f.use_reg) {
        const float pr = xv[0] * xv[1];
        sgn = pr > 0.0f ? 1.0f : (pr < 0.0f ? -1.0f : 0.0f);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float v = xv[k];
        if (f.act && (f.all_elems || ((prm.elem_mask >> k) & 1u))) {
            float gt = gv[k];
            float rg = 0.0f;
            if (f.use_reg) rg = prm.reg_w * (xv[K == 2 ? 1 - k : 0] * sgn);
            if (f.use_triv) rg += prm.triv_w * (v * (v * (4.0f * v - 3.0f) + 1.0f));
            gt += prm.beta_reg * rg;
            if (f.use_bar) {
                // grad phi = -g/f per constraint (src/barrier_method.py:50): -1/x for f = -x, -1/(x-1) for f = x-1;
                // together (1 - 2x) / (x (x - 1)): one reciprocal.  The signs of the zeros are the reference's: on
                // either face the quotient is -inf (x = 0: 1/(-0), x = 1: -1/(+0)), so a pixel clipped onto a face
                // is thrown to +inf and clipped to 1, exactly like the numpy code.
                if (f.lower_only) gt += f.bt * __fdividef(-1.0f, v);
                else gt += f.bt * __fdividef(1.0f - 2.0f * v, v * (v - 1.0f));
            }
            v = v - prm.alpha * gt;
        }
        if (f.clip) {
            // np.clip semantics: a NaN stays a NaN (fmaxf / fminf would turn it into the bound and hide the reference's
            // own breakdown after a -1/0 barrier gradient, teeth_recon/experiments.py:737-749 at 1024^2)
            v = v < 0.0f ? 0.0f : v;
            if (!f.lower_only) v = v > 1.0f ? 1.0f : v;
        }
        xn[k] = v;
    }
    float pen = 0.0f;
    if (f.want_pen) {
        if (f.use_bar) {
            float ph = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // -log(x) - log(1-x) = -log(x (1-x)); +inf on (or outside) the boundary like phi[f >= 0] = inf
                const float arg = f.lower_only ? xn[k] : xn[k] * (1.0f - xn[k]);
                const bool inside = f.lower_only ? (xn[k] > 0.0f) : (xn[k] > 0.0f && xn[k] < 1.0f);
                ph += inside ? -(f.fast_log ? __logf(arg) : logf(arg)) : INFINITY;
            }
            pen = f.bt * ph;
        }
        if (f.use_reg) pen += prm.beta_reg * (prm.reg_w * fabsf(xn[0] * xn[1]));
    }
    return pen;
}

// ---- K3 fused into K2: the backprojector's epilogue takes the gradient step itself -------------------------------
// A group of V interleaved images holds V/K slices x K elements, so a pixel's thread has the whole gradient of its
// slices in registers: x' = upd_pixel(x, gscale * W^T(q mu)) is written back (plain layout) and straight into K1's
// packed copies VN / VT for the next forward projection -- no gradient array, no K3 pass, no re-pack (a white-beam
// barrier iteration is K2 -> K1 -> one reduction: three launches where the separate entry points take seven).
template <int K>
struct BpEpiBarrierUpdate {
    static constexpr bool kHasSums = true;
    static constexpr int kMult = K;
    float *x;                      // (S, K, R, C), updated in place
    wt_update_params prm;
    const unsigned char *active;   // (S) or null
    double *partial;               // (S, nslots) penalty partial sums, or null
    float *vn, *vt;                // K1's packed copies, refreshed in place
    float gscale;                  // gradient = gscale * W^T(.)   (-1 for the white-beam data term)
    template <int V>
    __device__ __forceinline__ void operator()(const BpParams &p, int g, int row, int col, const float (&val)[V],
                                               float *sums) const {
        static_assert(V % K == 0 && V / K <= BP_MAX_SUMS, "group must hold whole slices");
        constexpr int SPG = V / K;
        typedef typename VecOf<V>::T VecT;
        const size_t img = (size_t)p.R * p.C, px = (size_t)row * p.C + col;
        float nv[V];
#pragma unroll
        for (int sl = 0; sl < SPG; ++sl) {
            const int slice = g * SPG + sl;
            const bool real = slice * K < p.nimg;
            UpdFlags f;
            const bool sl_on = real && (!active || active[slice]);
            f.act = sl_on && prm.mode != WT_UPD_CLIP_ONLY;
            f.clip = sl_on && prm.mode != WT_UPD_STEP_ONLY;
            f.use_reg = (K == 2) && prm.reg_w != 0.0f;
            f.use_bar = prm.t != 0.0f;
            f.lower_only = prm.lower_only != 0;
            f.use_triv = prm.triv_w != 0.0f;
            f.all_elems = (prm.elem_mask & ((1u << K) - 1u)) == ((1u << K) - 1u);
            f.fast_log = f.use_reg && f.use_bar && !f.lower_only && !f.use_triv && f.all_elems && prm.mode == WT_UPD_STEP_CLIP;
            f.want_pen = partial != nullptr && real;
            f.bt = prm.beta_reg * prm.t;
            float xv[K], gv[K], xn[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                xv[k] = real ? x[((size_t)slice * K + k) * img + px] : 0.0f;
                gv[k] = gscale * val[sl * K + k];
            }
            const float pen = upd_pixel<K>(prm, f, xv, gv, xn);
            if (f.want_pen) sums[sl] += pen;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                nv[sl * K + k] = real ? xn[k] : 0.0f;
                if (f.act || f.clip) x[((size_t)slice * K + k) * img + px] = xn[k];
            }
        }
        reinterpret_cast<VecT *>(vn)[((size_t)g * p.R + row) * fp_pitch(p.C) + VPAD + col] = pack<V>(nv);
        reinterpret_cast<VecT *>(vt)[((size_t)g * p.C + col) * fp_pitch(p.R) + VPAD + row] = pack<V>(nv);
    }
    template <int V>
    __device__ __forceinline__ void finish(const BpParams &p, int g, int slot, int nslots, const float *sums) const {
        if (!partial) return;
        constexpr int SPG = V / K;
#pragma unroll
        for (int sl = 0; sl < SPG; ++sl)
            if ((g * SPG + sl) * K < p.nimg) partial[(size_t)(g * SPG + sl) * nslots + slot] = (double)sums[sl];
    }
};

// loss[s] = 1/2 sum_b lpart[s][b] (float partials of K1's epilogue) and pen[s] = sum_b ppart[s][b] (float64 partials of
// K2's update epilogue), one launch: grid (S, 2), one warp each, fixed order
__global__ void reduce_step_kernel(const float *__restrict__ lpart, int nlb, double *loss, const double *__restrict__ ppart,
                                   int npb, double *pen) {
    const int s = blockIdx.x;
    double acc = 0.0;
    if (blockIdx.y == 0) {
        if (!loss) return;
        for (int b = threadIdx.x; b < nlb; b += 32) acc += (double)lpart[(size_t)s * nlb + b];
        acc = warp_sum(acc);
        if (threadIdx.x == 0) loss[s] = 0.5 * acc;
    } else {
        if (!pen || !ppart) return;
        for (int b = threadIdx.x; b < npb; b += 32) acc += ppart[(size_t)s * npb + b];
        acc = warp_sum(acc);
        if (threadIdx.x == 0) pen[s] = acc;
    }
}

// VEC = 4: every thread owns 4 consecutive pixels (16-byte loads/stores), VEC = 1: scalar fallback for
// pixel counts that are not a multiple of 4.  grid = (blocks per slice, slices).
// FAST: the flag set of the white-beam box-barrier iteration (K = 2 with the |c0 c1| regulariser, both barrier
// faces, all elements, step + clip) is fixed at compile time so that the per-element code is branch-free.
template <int K, int VEC, bool FAST>
__global__ void __launch_bounds__(256) barrier_update_kernel(float *__restrict__ x, const float *__restrict__ grad,
                                                             size_t npix, wt_update_params prm,
                                                             const unsigned char *__restrict__ active, double *partial) {
    typedef typename VecOf<VEC>::T VecT;
    const size_t slice = blockIdx.y;
    float *xs = x + slice * K * npix;
    const float *gs = grad ? grad + slice * K * npix : nullptr;
    UpdFlags f;
    const bool sl_on = !active || active[slice];
    f.act = FAST ? sl_on : (sl_on && prm.mode != WT_UPD_CLIP_ONLY);
    f.clip = FAST ? sl_on : (sl_on && prm.mode != WT_UPD_STEP_ONLY);
    f.use_reg = FAST ? true : ((K == 2) && prm.reg_w != 0.0f);
    f.use_bar = FAST ? true : (prm.t != 0.0f);
    f.lower_only = FAST ? false : (prm.lower_only != 0);
    f.use_triv = FAST ? false : (prm.triv_w != 0.0f);
    f.all_elems = FAST ? true : ((prm.elem_mask & ((1u << K) - 1u)) == ((1u << K) - 1u));
    f.fast_log = FAST;
    f.want_pen = partial != nullptr;
    f.bt = prm.beta_reg * prm.t;
    const bool writes = f.act || f.clip;
    float pen = 0.0f;
    const size_t nvec = npix / VEC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        float xv[K][VEC], gv[K][VEC], xo[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            unpack<VEC>(reinterpret_cast<const VecT *>(xs + k * npix)[i], xv[k]);
            if (f.act) {
                unpack<VEC>(__ldg(reinterpret_cast<const VecT *>(gs + k * npix) + i), gv[k]);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) gv[k][j] = 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float a[K], g[K], o[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { a[k] = xv[k][j]; g[k] = gv[k][j]; }
            pen += upd_pixel<K>(prm, f, a, g, o);
#pragma unroll
            for (int k = 0; k < K; ++k) xo[k][j] = o[k];
        }
        if (writes) {
#pragma unroll
            for (int k = 0; k < K; ++k) reinterpret_cast<VecT *>(xs + k * npix)[i] = pack<VEC>(xo[k]);
        }
    }
    if (partial) {
        __shared__ double red[8];
        double pd = warp_sum((double)pen);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pd;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            partial[slice * gridDim.x + blockIdx.x] = s;
        }
    }
}

}  // namespace wt
