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
    return 0;
}
