// What does ONE thread-per-row Gauss-Seidel slice update cost on the dependency path?  One warp, one SELL-32 slice whose entries are
// L2-resident, x in shared or global memory; cycles (clock64, min of 50 repetitions) of the pieces of the ordered thread-per-row kernels:
// descriptor + entries (prologue), gathers + in-order chain + IEEE division + store (gs_finish_sell).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Iinclude -o tools/ubench5 tools/ubench5.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../amg_b200/csrc/kernels.cuh"
using namespace amgb200;

template <int MODE, bool LAT>      // LAT: branch-free form (kernels.cuh); 0: x in shared, generic finish | 1: x in global (L1) | 2: x in global at L2 (COH) | 3: x in shared, single-chunk finish
__global__ void __launch_bounds__(32) slice_kernel(DMat A, const double *b, double *xg, int n, long long *out) {
    extern __shared__ double xs[];
    const int lane = threadIdx.x;
    for (int i = lane; i < n; i += 32) xs[i] = xg[i];
    __syncwarp();
    double *x = (MODE == 0 || MODE == 3) ? xs : xg;
    long long best[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30};
    double sink = 0.0;
    for (int rep = 0; rep < 50; ++rep) {
        const int s = rep % A.nitems;
        __syncwarp();
        const long long c0 = clock64();
        SellItem<20> ws;
        ws.prologue(A, s, lane, b);
        // consume one loaded value so that the clock read waits for the loads
        double probe = ws.a[0] + ws.bk + (double)ws.j[0];
        if (probe == 1.2345e300) sink += 1;
        __syncwarp();
        const long long c1 = clock64();
        if (MODE == 3) gs_finish_sell_one<false, 20, LAT>(ws, x);
        else if (MODE == 2) gs_finish_sell<true, 20, LAT>(ws, x);
        else gs_finish_sell<false, 20, LAT>(ws, x);
        __syncwarp();
        __threadfence_block();
        const double back = x[ws.k < ws.r1 ? ws.k : 0];
        if (back == 1.2345e300) sink += 1;
        const long long c2 = clock64();
        // the quotient alone
        double q = gs_quotient(back + 1.0, ws.a[0] + 3.0, 0);
        if (q == 1.2345e300) sink += 1;
        const long long c3 = clock64();
        best[0] = min(best[0], c1 - c0); best[1] = min(best[1], c2 - c1); best[2] = min(best[2], c3 - c2); best[3] = min(best[3], c3 - c0);
    }
    if (lane == 0) { for (int i = 0; i < 4; ++i) out[i] = best[i]; out[4] = (long long)sink; }
}

int main() {
    const int n = 8192, ns = 64;
    for (int width : {7, 9, 13, 19}) {
        std::vector<int> slice_row(ns + 1); std::vector<long long> slice_ptr(ns + 1);
        std::vector<int> col; std::vector<double> val;
        for (int s = 0; s <= ns; ++s) { slice_row[s] = 32 * s; slice_ptr[s] = (long long)s * 32 * width; }
        col.resize((size_t)ns * 32 * width); val.resize(col.size());
        srand(1);
        for (int s = 0; s < ns; ++s) for (int e = 0; e < width; ++e) for (int r = 0; r < 32; ++r) {
            const int k = 32 * s + r;
            col[(size_t)s * 32 * width + 32 * e + r] = e == 0 ? k : (e == width - 1 && (r & 3) == 0 ? -1 : rand() % n);     // diagonal first (Galerkin order), some padding
            val[(size_t)s * 32 * width + 32 * e + r] = e == 0 ? 4.0 : -1.0 / (1 + rand() % 7);
        }
        DMat A{};
        A.kind = 0; A.nrows = n; A.ncols = n; A.nitems = ns; A.max_row = width; A.recip = 0;
        int *d_sr, *d_col; long long *d_sp; double *d_val, *d_b, *d_x; long long *d_out;
        cudaMalloc(&d_sr, (ns + 1) * 4); cudaMalloc(&d_sp, (ns + 1) * 8); cudaMalloc(&d_col, col.size() * 4); cudaMalloc(&d_val, val.size() * 8);
        cudaMalloc(&d_b, n * 8); cudaMalloc(&d_x, n * 8); cudaMalloc(&d_out, 64);
        cudaMemcpy(d_sr, slice_row.data(), (ns + 1) * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_sp, slice_ptr.data(), (ns + 1) * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_col, col.data(), col.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_val, val.data(), val.size() * 8, cudaMemcpyHostToDevice);
        std::vector<double> ones(n, 1.0);
        cudaMemcpy(d_b, ones.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_x, ones.data(), n * 8, cudaMemcpyHostToDevice);
        A.slice_row = d_sr; A.slice_ptr = d_sp; A.col = d_col; A.val = d_val;
        long long h[5];
        const char *names[4] = {"x in shared memory, gs_finish_sell", "x in global memory (L1)", "x in global memory (L2, COH)", "x in shared memory, gs_finish_sell_one"};
        for (int lat = 0; lat < 2; ++lat) for (int mode = 0; mode < 4; ++mode) {
            if (mode == 0) { if (lat) { cudaFuncSetAttribute(slice_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<0, true><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } else { cudaFuncSetAttribute(slice_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<0, false><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } }
            if (mode == 1) { if (lat) { cudaFuncSetAttribute(slice_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<1, true><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } else { cudaFuncSetAttribute(slice_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<1, false><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } }
            if (mode == 2) { if (lat) { cudaFuncSetAttribute(slice_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<2, true><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } else { cudaFuncSetAttribute(slice_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<2, false><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } }
            if (mode == 3) { if (lat) { cudaFuncSetAttribute(slice_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<3, true><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } else { cudaFuncSetAttribute(slice_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8); slice_kernel<3, false><<<1, 32, n * 8>>>(A, d_b, d_x, n, d_out); } }
            cudaMemcpy(h, d_out, 40, cudaMemcpyDeviceToHost);
            printf("width %2d  %s  %-40s prologue (descriptor + entries from L2) %5lld   finish (gathers + chain + division + store) %5lld   division alone %4lld   total %5lld cycles  (%s)\n",
                   width, lat ? "branch-free" : "per-entry if", names[mode], h[0], h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(d_sr); cudaFree(d_sp); cudaFree(d_col); cudaFree(d_val); cudaFree(d_b); cudaFree(d_x); cudaFree(d_out);
    }
    return 0;
}
