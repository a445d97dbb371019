// Micro-benchmark of the product-folding chain exactly as the kernels use it (chain_fold from kernels.cuh)
// against variants, at several warps per CTA.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#include "../amg_b200/csrc/kernels.cuh"
using namespace amgb200;

// variant B: loads pinned by volatile asm (no __syncwarp)
__device__ __forceinline__ double2 lds2(const double2 *p) {
    double2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double fold_b(double t, const double2 *sp2, int cnt) {
#define FOLD(v) _Pragma("unroll") for (int u = 0; u < 4; ++u) { t = __dsub_rn(t, v[u].x); t = __dsub_rn(t, v[u].y); }
#define LOAD(v, q) _Pragma("unroll") for (int u = 0; u < 4; ++u) v[u] = lds2(sp2 + ((q) >> 1) + u);
    double2 va[4], vb[4];
    int q = 0;
    LOAD(va, 0)
#pragma unroll 1
    for (; q + 32 <= cnt; q += 32) {
        LOAD(vb, q + 8)  FOLD(va)
        LOAD(va, q + 16) FOLD(vb)
        LOAD(vb, q + 24) FOLD(va)
        LOAD(va, q + 32) FOLD(vb)
    }
#pragma unroll 1
    for (; q < cnt; q += 8) {
        LOAD(vb, q + 8)  FOLD(va)
#pragma unroll
        for (int u = 0; u < 4; ++u) va[u] = vb[u];
    }
#undef FOLD
#undef LOAD
    return t;
}
// variant C: simple loop, plain loads (let ptxas schedule)
__device__ __forceinline__ double fold_c(double t, const double2 *sp2, int cnt) {
#pragma unroll 4
    for (int q = 0; q < cnt; q += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { double2 v = sp2[(q >> 1) + u]; t = __dsub_rn(t, v.x); t = __dsub_rn(t, v.y); }
    }
    return t;
}
template <int V>
__global__ void fold_bench(double *out, int cnt, int reps) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sp = sm + warp * 1024;
    for (int i = lane; i < 1024; i += 32) sp[i] = i < cnt ? 1e-9 * (i + 1) : 0.0;
    __syncthreads();
    double t = 1.0;
    long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (V == 0) t = chain_fold<true>(t, (const double2 *)sp, cnt);
        if (V == 1) t = fold_b(t, (const double2 *)sp, cnt);
        if (V == 2) t = fold_c(t, (const double2 *)sp, cnt);
    }
    long long c1 = clock64();
    if (threadIdx.x == 0) { out[0] = t; out[1] = (double)(c1 - c0) / ((double)reps * cnt); }
}
int main() {
    double *d; cudaMalloc(&d, 64);
    const char *nm[3] = {"chain_fold (syncwarp fences)", "volatile asm loads", "plain loop"};
    for (int v = 0; v < 3; ++v)
        for (int cnt : {24, 88, 128, 344, 1000})
            for (int warps : {1, 4, 8, 12, 16}) {
                const size_t sh = warps * 1024 * 8;
                auto k = v == 0 ? fold_bench<0> : v == 1 ? fold_bench<1> : fold_bench<2>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                k<<<1, warps * 32, sh>>>(d, cnt, 64);
                cudaError_t e = cudaDeviceSynchronize();
                double r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
                printf("%-30s cnt %4d warps %2d : %.2f cycles/term (incl. call overhead) %s\n", nm[v], cnt, warps, r[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
