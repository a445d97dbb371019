// How do K independent loads of one thread/warp overlap, by load flavour?  (L2-resident data, one warp)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int K>
__global__ void lat(const uint4 *p, long long *out, int stride) {
    const int lane = threadIdx.x;
    uint4 r[K];
    const uint4 *q = p + lane;
    // warm L2 (not L1): touch with cg
    for (int u = 0; u < K; ++u) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(q + (size_t)u * stride));
    unsigned acc = 0;
    for (int u = 0; u < K; ++u) acc += r[u].x;
    __syncwarp();
    long long best = 1 << 30;
    for (int rep = 0; rep < 20; ++rep) {
        long long c0 = clock64();
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const uint4 *a = q + (size_t)u * stride;
            if (MODE == 0) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(a) : "memory");
            if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(a) : "memory");
            if (MODE == 2) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(a) : "memory");
            if (MODE == 3) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(a) : "memory");
            if (MODE == 4) asm volatile("ld.weak.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(a) : "memory");
            if (MODE == 5) asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u].x), "=r"(r[u].y), "=r"(r[u].z), "=r"(r[u].w) : "l"(a) : "memory");
        }
#pragma unroll
        for (int u = 0; u < K; ++u) acc += r[u].x + r[u].w;
        long long c1 = clock64();
        if (c1 - c0 < best) best = c1 - c0;
        __syncwarp();
    }
    if (lane == 0) { out[0] = best; out[1] = acc; }
}
template <int MODE>
void run(const char *name, const uint4 *p, long long *out) {
    long long h[2];
    printf("%-34s", name);
    lat<MODE, 1><<<1, 32>>>(p, out, 64); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("  K=1 %6lld", h[0]);
    lat<MODE, 4><<<1, 32>>>(p, out, 64); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("  K=4 %6lld", h[0]);
    lat<MODE, 10><<<1, 32>>>(p, out, 64); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("  K=10 %6lld", h[0]);
    lat<MODE, 20><<<1, 32>>>(p, out, 64); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("  K=20 %6lld cycles  (%s)\n", h[0], cudaGetErrorString(cudaGetLastError()));
}
int main() {
    uint4 *p; long long *out;
    cudaMalloc(&p, 64 << 20); cudaMemset(p, 1, 64 << 20); cudaMalloc(&out, 16);
    run<0>("ld.global.cg", p, out);
    run<1>("ld.relaxed.gpu", p, out);
    run<2>("ld.volatile", p, out);
    run<3>("ld.global.cv", p, out);
    run<4>("ld.weak.global.cg", p, out);
    run<5>("ld.relaxed.gpu L1::no_allocate", p, out);
    return 0;
}
