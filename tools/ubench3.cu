// Micro-benchmarks for the sync-free (data-flow) Gauss-Seidel design: how long does one dependency hop between
// two SMs take when the consumer polls (a) a packed 16-byte {x, version} record, (b) a flag written after a fence.
// Also integer-chain latencies for the fixed-point fold.   nvcc -O3 -arch=sm_100a -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct __align__(16) Rec { unsigned lo, v0, hi, v1; };

__device__ __forceinline__ void st_rec(Rec *p, double x, unsigned ver) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(ver), "r"((unsigned)(b >> 32)), "r"(ver) : "memory");
}
__device__ __forceinline__ bool ld_rec(const Rec *p, unsigned need, double &x, unsigned *torn) {
    unsigned lo, v0, hi, v1;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(v0), "=r"(hi), "=r"(v1) : "l"(p) : "memory");
    if (v0 != v1 && torn) atomicAdd(torn, 1u);
    if (v0 == v1 && (int)(v0 - need) >= 0) { x = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo)); return true; }
    return false;
}

// ring of G CTAs (one warp each, lane 0 active): hop k is done by CTA k % G; it needs the record written by hop k-1
__global__ void hop_packed(Rec *rec, int G, int hops, unsigned *torn, unsigned long long *out) {
    if (threadIdx.x) return;
    const int me = blockIdx.x;
    unsigned long long t0 = 0;
    if (me == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = me; k < hops; k += G) {
        double x = 1.0;
        if (k > 0) { const Rec *src = rec + (k - 1) % G; while (!ld_rec(src, (unsigned)k, x, torn)) {} }
        // the consumer checks that the value belongs to the version (catches torn or stale data)
        if (k > 0 && x != (double)(k - 1) * 0.5 + 1.0) atomicAdd(torn + 1, 1u);
        st_rec(rec + me, (double)k * 0.5 + 1.0, (unsigned)(k + 1));
    }
    if (me == (hops - 1) % G) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[1] = t1;
    }
    if (me == 0) out[0] = t0;
}

// same ring, classic protocol: x, __threadfence, flag ; consumer: poll flag (acquire), then load x
__global__ void hop_flag(double *xs, unsigned *flag, int G, int hops, unsigned *torn, unsigned long long *out) {
    if (threadIdx.x) return;
    const int me = blockIdx.x;
    unsigned long long t0 = 0;
    if (me == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = me; k < hops; k += G) {
        double x = 1.0;
        if (k > 0) {
            const int s = (k - 1) % G;
            unsigned f;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag + s * 32) : "memory"); } while ((int)(f - (unsigned)k) < 0);
            x = __ldcg(xs + s * 16);
            if (x != (double)(k - 1) * 0.5 + 1.0) atomicAdd(torn + 1, 1u);
        }
        __stcg(xs + me * 16, (double)k * 0.5 + 1.0);
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag + me * 32), "r"((unsigned)(k + 1)) : "memory");
    }
    if (me == (hops - 1) % G) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[1] = t1;
    }
    if (me == 0) out[0] = t0;
}

// fan-in variant of the packed protocol: every hop waits for F records written by the previous F hops' CTAs ... here: all
// lanes of the warp poll one record each (lane l polls the record of hop k-1-l when it exists), like a row with 32 late entries
__global__ void hop_packed_fan(Rec *rec, int G, int hops, int fan, unsigned *torn, unsigned long long *out) {
    const int me = blockIdx.x, lane = threadIdx.x;
    unsigned long long t0 = 0;
    if (me == 0 && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = me; k < hops; k += G) {
        double x = 1.0;
        const int dep = k - 1 - lane;
        if (lane < fan && dep >= 0) { const Rec *src = rec + dep % G; while (!ld_rec(src, (unsigned)(dep + 1), x, torn)) {} }
        double s = x;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) st_rec(rec + me, s * 1e-30 + (double)k * 0.5 + 1.0, (unsigned)(k + 1));
        __syncwarp();
    }
    if (me == (hops - 1) % G && lane == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[1] = t1;
    }
    if (me == 0 && lane == 0) out[0] = t0;
}

// dependent integer chains: 64-bit add, and the tie step (N - I) & ~tie
__global__ void iadd_chain(long long *out, const long long *in, int n) {
    long long N = in[0];
    const long long I = in[1];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) { N -= I; N -= (I ^ N) & 1; N -= I; N -= (I ^ N) & 1; }
    long long c1 = clock64();
    out[0] = N; out[1] = (c1 - c0);
}
__global__ void iadd_chain_plain(long long *out, const long long *in, int n) {
    long long N = in[0];
    const long long I = in[1], J = in[2];
    long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { N -= I; N += J; N -= I; N += J; N -= I; N += J; N -= I; N += J; }
    long long c1 = clock64();
    out[0] = N; out[1] = (c1 - c0);
}

int main() {
    Rec *rec; double *xs; unsigned *flag, *torn; unsigned long long *out;
    cudaMalloc(&rec, 4096 * sizeof(Rec)); cudaMalloc(&xs, 4096 * 16 * 8); cudaMalloc(&flag, 4096 * 32 * 4);
    cudaMalloc(&torn, 8); cudaMalloc(&out, 16);
    const int hops = 200000;
    for (int G : {2, 8, 64, 148}) {
        unsigned long long h[2]; unsigned ht[2];
        cudaMemset(rec, 0, 4096 * sizeof(Rec)); cudaMemset(torn, 0, 8);
        hop_packed<<<G, 32>>>(rec, G, hops, torn, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); cudaMemcpy(ht, torn, 8, cudaMemcpyDeviceToHost);
        printf("packed 16-byte record, ring of %3d CTAs: %.1f ns per hop   (torn reads %u, wrong values %u)  %s\n", G, (double)(h[1] - h[0]) / hops, ht[0], ht[1], cudaGetErrorString(cudaGetLastError()));
        cudaMemset(flag, 0, 4096 * 32 * 4); cudaMemset(torn, 0, 8);
        hop_flag<<<G, 32>>>(xs, flag, G, hops, torn, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); cudaMemcpy(ht, torn, 8, cudaMemcpyDeviceToHost);
        printf("x + release flag,      ring of %3d CTAs: %.1f ns per hop   (wrong values %u)  %s\n", G, (double)(h[1] - h[0]) / hops, ht[1], cudaGetErrorString(cudaGetLastError()));
        for (int fan : {8, 32}) {
            cudaMemset(rec, 0, 4096 * sizeof(Rec)); cudaMemset(torn, 0, 8);
            hop_packed_fan<<<G, 32>>>(rec, G, hops, fan, torn, out);
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); cudaMemcpy(ht, torn, 8, cudaMemcpyDeviceToHost);
            printf("packed, fan-in %2d,      ring of %3d CTAs: %.1f ns per hop   (torn reads %u)  %s\n", fan, G, (double)(h[1] - h[0]) / hops, ht[0], cudaGetErrorString(cudaGetLastError()));
        }
    }
    long long *din, *dout, hin[3] = {(1LL << 52) + 12345, 3, 5}, hout[2];
    cudaMalloc(&din, 24); cudaMalloc(&dout, 16);
    cudaMemcpy(din, hin, 24, cudaMemcpyHostToDevice);
    iadd_chain<<<1, 32>>>(dout, din, 10000); cudaMemcpy(hout, dout, 16, cudaMemcpyDeviceToHost);
    printf("int64 chain with parity step: %.2f cycles per (sub + tie-adjust) pair\n", (double)hout[1] / 20000.0);
    iadd_chain_plain<<<1, 32>>>(dout, din, 10000); cudaMemcpy(hout, dout, 16, cudaMemcpyDeviceToHost);
    printf("int64 dependent add chain: %.2f cycles per add\n", (double)hout[1] / 80000.0);
    return 0;
}
