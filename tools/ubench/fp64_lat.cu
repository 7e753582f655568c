// FP64 pipe micro-benchmark: cycles per warp-wide DFMA for ILP independent chains per thread and W warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm, double *out, long long *cyc) {
    const int iters = 4096;
    k<ILP><<<148, warps_per_sm * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<ILP><<<148, warps_per_sm * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp_instr = (double)h / (iters * ILP);
    const double wps = warps_per_sm / 4.0;
    printf("ILP %d warps/SMSP %.0f: %.2f cycles per DFMA per warp; SMSP issues one DFMA per %.2f cycles\n", ILP, wps, per_warp_instr, per_warp_instr / wps);
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 16, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
    return 0;
}
