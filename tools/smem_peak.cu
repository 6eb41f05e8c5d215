// Measures the on-chip ceiling that bounds K1/K2: conflict-free shared-memory load bandwidth (LDS.128 and
// LDS.32) summed over all SMs.  Prints one JSON line; the result is committed as profiles/ONCHIP_PEAKS.json
// and is the denominator of the "fraction of the shared-memory ceiling" quoted in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/smem_peak tools/smem_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void __launch_bounds__(512) lds_kernel(float *out, int iters) {
    extern __shared__ __align__(16) unsigned char raw[];
    T *buf = reinterpret_cast<T *>(raw);
    const int n = 32768 / sizeof(T);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = (float)i;
        T t;
        float *p = reinterpret_cast<float *>(&t);
        for (int k = 0; k < (int)(sizeof(T) / 4); ++k) p[k] = v;
        buf[i] = t;
    }
    __syncthreads();
    float acc[sizeof(T) / 4];
    for (int k = 0; k < (int)(sizeof(T) / 4); ++k) acc[k] = 0.f;
    int idx = threadIdx.x;  // consecutive lanes -> consecutive T: conflict-free
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            T t = buf[(idx + u * (n / 8)) & (n - 1)];   // 8 distinct addresses per iteration
            const float *p = reinterpret_cast<const float *>(&t);
#pragma unroll
            for (int k = 0; k < (int)(sizeof(T) / 4); ++k) acc[k] += p[k];
        }
        idx = (idx + 37) & (n - 1);
    }
    float s = 0.f;
    for (int k = 0; k < (int)(sizeof(T) / 4); ++k) s += acc[k];
    if (s == -1.2345f) out[0] = s;
}

template <typename T>
static double run(int sms, int iters) {
    float *out;
    cudaMalloc(&out, 4);
    const int blocks = sms * 4;
    cudaFuncSetAttribute(lds_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    lds_kernel<T><<<blocks, 512, 32768>>>(out, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        lds_kernel<T><<<blocks, 512, 32768>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)blocks * 512 * iters * 8 * sizeof(T);
        const double gbs = bytes / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double v128 = run<float4>(p.multiProcessorCount, 4096);
    const double v32 = run<float>(p.multiProcessorCount, 4096);
    // one update of K1/K2 = 2 taps; with V = 4 interleaved images a tap is 16 B for 4 updates -> 8 B per update
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"lds128_gbs\": %.1f, \"lds32_gbs\": %.1f, \"gups_ceiling_v4\": %.1f, "
           "\"nominal_gups_ceiling_v4\": %.1f, \"how\": \"conflict-free LDS from a 32 KB buffer, 4 CTAs x 512 threads per SM, "
           "best of 5, CUDA events; nominal = 128 B/clk/SM x SMs x max SM clock / 8 B per update\"}\n",
           p.name, p.multiProcessorCount, v128, v32, v128 / 8.0, 128.0 * p.multiProcessorCount * (p.clockRate * 1e3) / 8.0 / 1e9);
    return 0;
}
