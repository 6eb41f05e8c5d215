// K1: ray-driven Joseph forward projector (ASTRA 'linear' projector semantics),
// V images interleaved per position so that one 4V-byte load feeds V updates.
//
// Work decomposition: thread <-> ray (angle a, detector d); a warp holds 32 adjacent detectors of one
// angle, a CTA FP_DT detectors x FP_AT consecutive angles (consecutive angles walk almost the same
// pixels, so their warps share L1 lines).  "vertical" angles march over rows of the row-major packed
// volume, "horizontal" angles over rows of the transposed packed copy -- both are the same loop, and
// adjacent lanes always read adjacent positions of one line (coalesced).
#pragma once
#include "common.cuh"

namespace wt {

constexpr int FP_DT = 64;  // detectors per CTA
constexpr int FP_AT = 4;   // angles per CTA

// ---- packing: plain (n, R, C) -> VN (G, R, C+2*VPAD, V) and VT (G, C, R+2*VPAD, V), zero padded -------
template <int V>
__global__ void __launch_bounds__(256) pack_vol_kernel(const float *__restrict__ vol, float *__restrict__ vn,
                                                       float *__restrict__ vt, int R, int C) {
    typedef typename VecOf<V>::T VecT;
    __shared__ float tile[V][32][33];
    const int g = blockIdx.z;
    const int r0 = blockIdx.y * 32 - VPAD, c0 = blockIdx.x * 32 - VPAD;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Cp = C + 2 * VPAD, Rp = R + 2 * VPAD;
    const size_t img = (size_t)R * C;
    const float *src = vol + (size_t)g * V * img;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rr = ty + 8 * k;
        const int r = r0 + rr, c = c0 + tx;
        float v[V];
        const bool in = (r >= 0 && r < R && c >= 0 && c < C);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = in ? __ldg(src + i * img + (size_t)r * C + c) : 0.0f;
            tile[i][rr][tx] = v[i];
        }
        if (r >= 0 && r < R && c + VPAD < Cp)
            reinterpret_cast<VecT *>(vn)[((size_t)g * R + r) * Cp + (c + VPAD)] = pack<V>(v);
    }
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

template <int V>
inline void launch_pack_vol(const float *vol, float *vn, float *vt, int R, int C, int groups, cudaStream_t st) {
    dim3 grid((C + 2 * VPAD + 31) / 32, (R + 2 * VPAD + 31) / 32, groups);
    pack_vol_kernel<V><<<grid, dim3(32, 8), 0, st>>>(vol, vn, vt, R, C);
}

// ---- projector ------------------------------------------------------------------------------------
struct FpParams {
    const float *vn;
    const float *vt;
    const FpAngle *ang;
    int R, C, D, A;
    int a_begin, a_end;
};

// plain store: sino (n, A, D)
struct FpEpiStore {
    float *sino;
    template <int V>
    __device__ __forceinline__ void operator()(const FpParams &p, int g, int a, int d, bool valid,
                                               const float (&val)[V]) const {
        if (!valid) return;
        const size_t plane = (size_t)p.A * p.D;
#pragma unroll
        for (int i = 0; i < V; ++i) sino[((size_t)g * V + i) * plane + (size_t)a * p.D + d] = val[i];
    }
};

template <int V>
__device__ __forceinline__ void fp_ray(const FpParams &p, const FpAngle &an, int g, int d, float (&val)[V]) {
    typedef typename VecOf<V>::T VecT;
    const int nlines = an.vertical ? p.R : p.C;
    const int npos = an.vertical ? p.C : p.R;
    const int pitch = npos + 2 * VPAD;
    const VecT *base = reinterpret_cast<const VecT *>(an.vertical ? p.vn : p.vt) + (size_t)g * nlines * pitch + VPAD;
    const float c0 = (float)(an.c0_base + ((double)d + 0.5) * an.c0_step);
    const float delta = an.delta;

    // lines on which floor(c) can be inside [-1, npos-1]; the estimate may be off by one line on each
    // side, which the zero padding absorbs (|delta| <= 1, VPAD = 4).
    int l_lo, l_hi;
    if (fabsf(delta) * (float)nlines < 0.5f) {
        const bool hit = (c0 > -2.0f) && (c0 < (float)npos + 1.0f);
        l_lo = 0;
        l_hi = hit ? nlines : 0;
    } else {
        const float inv = 1.0f / delta;
        const float t0 = (-1.0f - c0) * inv, t1 = ((float)npos - c0) * inv;
        const float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
        l_lo = (int)fmaxf(floorf(lo), 0.0f);
        l_hi = (int)fminf(ceilf(hi) + 1.0f, (float)nlines);
    }

    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.0f;

    const VecT *line = base + (size_t)l_lo * pitch;
    float lf = (float)l_lo;
#pragma unroll 4
    for (int l = l_lo; l < l_hi; ++l) {
        const float c = fmaf(lf, delta, c0);
        const float fl = floorf(c);
        const float off = c - fl;
        const int col = (int)fl;
        float v0[V], v1[V];
        unpack<V>(__ldg(line + col), v0);
        unpack<V>(__ldg(line + col + 1), v1);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += fmaf(off, v1[i] - v0[i], v0[i]);
        line += pitch;
        lf += 1.0f;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) val[i] = acc[i] * an.len;
}

template <int V, class Epi>
__global__ void __launch_bounds__(FP_DT *FP_AT) fp_kernel(FpParams p, Epi epi) {
    const int d = blockIdx.x * FP_DT + threadIdx.x;
    const int a = p.a_begin + blockIdx.y * FP_AT + threadIdx.y;
    const int g = blockIdx.z;
    const bool valid = (d < p.D) && (a < p.a_end);
    float val[V];
#pragma unroll
    for (int i = 0; i < V; ++i) val[i] = 0.0f;
    if (valid) {
        const FpAngle an = p.ang[a];
        fp_ray<V>(p, an, g, d, val);
    }
    epi.template operator()<V>(p, g, a, d, valid, val);
}

template <int V, class Epi>
inline void launch_fp(const FpParams &p, const Epi &epi, int groups, cudaStream_t st) {
    dim3 grid((p.D + FP_DT - 1) / FP_DT, (p.a_end - p.a_begin + FP_AT - 1) / FP_AT, groups);
    fp_kernel<V, Epi><<<grid, dim3(FP_DT, FP_AT), 0, st>>>(p, epi);
}

}  // namespace wt
