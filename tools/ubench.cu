// Micro-benchmarks for the latency constants that bound the in-order fp64 accumulation chain.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dadd_chain(double *out, double x, int n) {
    double t = out[0];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) { t = __dadd_rn(t, x); t = __dadd_rn(t, x); t = __dadd_rn(t, x); t = __dadd_rn(t, x); }
    long long c1 = clock64();
    out[0] = t; out[1] = (double)(c1 - c0) / (4.0 * n);
}
__global__ void dfma_chain(double *out, double x, int n) {
    double t = out[0];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) { t = __fma_rn(t, x, x); t = __fma_rn(t, x, x); t = __fma_rn(t, x, x); t = __fma_rn(t, x, x); }
    long long c1 = clock64();
    out[0] = t; out[1] = (double)(c1 - c0) / (4.0 * n);
}
__global__ void shfl_chain(double *out, double x, int n) {      // the EXACT warp-per-row inner loop
    double t = out[0], prod = x * (threadIdx.x + 1);
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int q = 0; q < 32; ++q) t = __dsub_rn(t, __shfl_sync(0xffffffffu, prod, q));
    }
    long long c1 = clock64();
    if (threadIdx.x == 0) { out[0] = t; out[1] = (double)(c1 - c0) / (32.0 * n); }
}
__global__ void lds_chain(double *out, double x, int n) {       // products staged in shared memory, broadcast reads
    __shared__ double prod[32 * 8];
    prod[threadIdx.x] = x * (threadIdx.x + 1);
    __syncthreads();
    double t = out[0];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) {
        const double2 *p2 = (const double2 *)(prod + (i & 7) * 32);
#pragma unroll
        for (int q = 0; q < 16; ++q) { double2 v = p2[q]; t = __dsub_rn(t, v.x); t = __dsub_rn(t, v.y); }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0) { out[0] = t; out[1] = (double)(c1 - c0) / (32.0 * n); }
}
__global__ void ldg_latency(const int *chain, int n, double *out) {   // pointer chase through L2-resident data
    int j = 0;
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) j = __ldcg(chain + j);
    long long c1 = clock64();
    out[0] = j; out[1] = (double)(c1 - c0) / n;
}
__global__ void ldg_l1_latency(const int *chain, int n, double *out) {
    int j = 0;
    for (int i = 0; i < 64; ++i) j = chain[j];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) j = chain[j];
    long long c1 = clock64();
    out[0] = j; out[1] = (double)(c1 - c0) / n;
}
int main() {
    double *d; cudaMalloc(&d, 64); double h[2] = {1.0, 0};
    auto run = [&](const char *name, auto f) { cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice); f(); cudaDeviceSynchronize(); double r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost); printf("%-28s %.2f cycles\n", name, r[1]); };
    run("dependent DADD", [&] { dadd_chain<<<1, 32>>>(d, 1e-9, 4096); });
    run("dependent DFMA", [&] { dfma_chain<<<1, 32>>>(d, 1e-9, 4096); });
    run("DSUB fed by SHFL (1 warp)", [&] { shfl_chain<<<1, 32>>>(d, 1e-9, 512); });
    run("DSUB fed by SHFL (8 warps)", [&] { shfl_chain<<<1, 256>>>(d, 1e-9, 512); });
    run("DSUB fed by LDS.128 (1 warp)", [&] { lds_chain<<<1, 32>>>(d, 1e-9, 512); });
    run("DSUB fed by LDS.128 (8 warps)", [&] { lds_chain<<<1, 256>>>(d, 1e-9, 512); });
    const int N = 1 << 20; int *hc = new int[N]; for (int i = 0; i < N; ++i) hc[i] = (int)(((long long)i * 7919 + 4099) % N);
    int *dc; cudaMalloc(&dc, N * 4); cudaMemcpy(dc, hc, N * 4, cudaMemcpyHostToDevice);
    run("L2 load latency (ld.cg chase)", [&] { ldg_latency<<<1, 1>>>(dc, 4096, d); });
    const int M = 1 << 12; for (int i = 0; i < M; ++i) hc[i] = (i * 33 + 7) % M; cudaMemcpy(dc, hc, M * 4, cudaMemcpyHostToDevice);
    run("L1 load latency (chase 16KB)", [&] { ldg_l1_latency<<<1, 1>>>(dc, 4096, d); });
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("SM clock attr %d kHz\n", clk);
    extern int main2();
    return main2();
}
// ---- synchronisation latencies ----------------------------------------------------------------
__global__ void __cluster_dims__(16, 1, 1) cluster_barrier_lat(double *out, double *buf, int n, int mode) {
    unsigned rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    long long c0 = clock64();
    double acc = 0;
    for (int i = 0; i < n; ++i) {
        if (mode >= 1 && threadIdx.x == 0) buf[rank * 32 + (i & 15)] = (double)i;         // a store before the release
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (mode >= 2) acc += __ldcg(buf + ((rank + 1) & 15) * 32 + (i & 15));            // a dependent L2 load after the acquire
    }
    long long c1 = clock64();
    if (threadIdx.x == 0 && rank == 0) { out[0] = acc; out[1] = (double)(c1 - c0) / n; }
}
__global__ void syncthreads_lat(double *out, int n) {
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long c1 = clock64();
    if (threadIdx.x == 0) out[1] = (double)(c1 - c0) / n;
}
__global__ void ddiv_lat(double *out, double x, int n) {
    double t = out[0] + 3.0;
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) { t = __ddiv_rn(t, x); t = __ddiv_rn(t, x); }
    long long c1 = clock64();
    out[0] = t; out[1] = (double)(c1 - c0) / (2.0 * n);
}
__global__ void grid_flag_lat(unsigned *flag, double *out, int n) {   // 2 CTAs ping-pong through an L2 flag (fence + atomic / poll)
    volatile unsigned *f = flag;
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) {
        if ((i & 1) == (int)blockIdx.x) { __threadfence(); atomicAdd(flag, 1u); }
        else { while (*f < (unsigned)(i + 1)) { } __threadfence(); }
    }
    long long c1 = clock64();
    if (blockIdx.x == 0) out[1] = (double)(c1 - c0) / n;
}
int main2() {
    double *d; cudaMalloc(&d, 64); double *buf; cudaMalloc(&buf, 16 * 32 * 8); cudaMemset(buf, 0, 16 * 32 * 8);
    auto rd = [&](const char *name) { cudaError_t e = cudaDeviceSynchronize(); double r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost); printf("%-44s %.1f cycles %s\n", name, r[1], e == cudaSuccess ? "" : cudaGetErrorString(e)); };
    cudaFuncSetAttribute(cluster_barrier_lat, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cluster_barrier_lat<<<16, 256>>>(d, buf, 2000, 0); rd("cluster barrier (16 CTAs x 256 thr)");
    cluster_barrier_lat<<<16, 256>>>(d, buf, 2000, 1); rd("cluster barrier + store before release");
    cluster_barrier_lat<<<16, 256>>>(d, buf, 2000, 2); rd("cluster barrier + store + L2 load after");
    cluster_barrier_lat<<<16, 32>>>(d, buf, 2000, 2); rd("same, 16 CTAs x 32 thr");
    syncthreads_lat<<<1, 512>>>(d, 4000); rd("__syncthreads (512 thr)");
    syncthreads_lat<<<1, 64>>>(d, 4000); rd("__syncthreads (64 thr)");
    ddiv_lat<<<1, 32>>>(d, 1.0000001, 2000); rd("dependent DDIV");
    unsigned *fl; cudaMalloc(&fl, 4); cudaMemset(fl, 0, 4);
    grid_flag_lat<<<2, 32>>>(fl, d, 2000); rd("L2 flag hop (fence+atomic -> poll), 2 CTAs");
    return 0;
}
