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
// the slot fold of the streaming smoother: S row slots per warp, each slot folds its own row
__global__ void slots_bench(double *out, int cnt, int reps, int S, int stagger) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *base = sm + 8 + warp * (S * 1032 + 64);
    for (int i = threadIdx.x; i < 8; i += blockDim.x) sm[i] = 0.0;
    for (int i = lane; i < S * 1032; i += 32) base[i] = 1e-9 * (i % 1000 + 1);
    __syncthreads();
    const int slot = lane / (32 / S);
    const double2 *sp2 = (const double2 *)(base + slot * (1024 + stagger));
    const int mycnt = cnt - 8 * (slot % 4);           // slots of slightly different length
    double t = 1.0;
    long long c0 = clock64();
    for (int r = 0; r < reps; ++r) t = chain_fold_slots(t, smem_u32(sp2), mycnt, cnt, smem_u32(sm));
    long long c1 = clock64();
    if (threadIdx.x == 0) { out[0] = t; out[1] = (double)(c1 - c0) / ((double)reps * cnt); }
}
// interference: warp 0 folds a chain; the other warps run the product pass of the streaming smoother (random x gathers,
// multiply, store) or a second chain
__global__ void interfere_bench(double *out, int cnt, int reps, int other_mode, const int *cols) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *xs = sm + 8;                 // 4096 doubles of "x"
    double *rows = xs + 4096;            // per warp 2048 doubles
    int *scol = (int *)(rows + 2048 * (blockDim.x >> 5));
    for (int i = threadIdx.x; i < 8; i += blockDim.x) sm[i] = 0.0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) { xs[i] = 1.0 + 1e-9 * i; scol[i] = cols[i]; }
    for (int i = threadIdx.x; i < 2048 * (blockDim.x >> 5); i += blockDim.x) rows[i] = 1e-9 * (i % 1000 + 1);
    __syncthreads();
    double *mine = rows + 2048 * warp;
    if (warp == 0 || other_mode == 2) {
        double t = 1.0;
        long long c0 = clock64();
        for (int r = 0; r < reps; ++r) t = chain_fold_slots(t, smem_u32(mine), cnt, cnt, smem_u32(sm));
        long long c1 = clock64();
        if (threadIdx.x == 0) { out[0] = t; out[1] = (double)(c1 - c0) / ((double)reps * cnt); }
        if (lane == 0 && warp) out[2 + warp] = t;
    } else if (other_mode == 1) {
        const unsigned val_a = smem_u32(mine), col_a = smem_u32(scol), x_a = smem_u32(xs);
        for (int r = 0; r < reps * 6; ++r)
            for (int p0 = 0; p0 < 512; p0 += 256) {
                int j[8]; double v[8], xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) j[u] = lds_s32(col_a + 4u * (unsigned)((p0 + u * 32 + lane + r * 37) & 4095));
#pragma unroll
                for (int u = 0; u < 8; ++u) { v[u] = lds_f64(val_a + 8u * (unsigned)(p0 + u * 32 + lane)); xv[u] = lds_f64(x_a + 8u * (unsigned)j[u]); }
#pragma unroll
                for (int u = 0; u < 8; ++u) sts_f64(val_a + 8u * (unsigned)(p0 + u * 32 + lane), __dmul_rn(v[u], xv[u]));
            }
    }
}
int main() {
    {
        double *d; cudaMalloc(&d, 256);
        int *hc = new int[4096]; for (int i = 0; i < 4096; ++i) hc[i] = (int)(((long long)i * 2654435761u) % 4096);
        int *dc; cudaMalloc(&dc, 4096 * 4); cudaMemcpy(dc, hc, 4096 * 4, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(interfere_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        const char *mn[3] = {"idle", "product pass", "own chain"};
        for (int mode : {0, 1, 2})
            for (int warps : {1, 2, 5, 9}) {
                const size_t sh = (8 + 4096 + 2048 * warps) * 8 + 4096 * 4;
                interfere_bench<<<1, warps * 32, sh>>>(d, 464, 64, mode, dc);
                cudaError_t e = cudaDeviceSynchronize();
                double r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
                printf("warp 0 chain, %d other warps doing [%s]: %.2f cycles/term %s\n", warps - 1, mn[mode], r[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    }
    {
        double *d; cudaMalloc(&d, 64);
        cudaFuncSetAttribute(slots_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int S : {1, 2, 8})
            for (int stagger : {0, 2, 8})
                for (int warps : {1, 2}) {
                    const size_t sh = (8 + warps * (S * 1032 + 64)) * 8;
                    slots_bench<<<1, warps * 32, sh>>>(d, 464, 64, S, stagger);
                    cudaError_t e = cudaDeviceSynchronize();
                    double r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
                    printf("chain_fold_slots S=%d stagger %d doubles, %d warps, cnt 464: %.2f cycles/term %s\n", S, stagger, warps, r[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    }
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
