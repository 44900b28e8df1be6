// dfma_probe.cu -- what the FP64 pipe of one B200 SM sub-partition needs to stay busy.
//   (1) dependent-issue latency of DFMA / DADD / DMUL (one warp, one chain, clock64 around 4096 dependent ops)
//   (2) DFMA throughput per SM as a function of resident warps per SM and independent chains per warp (ILP)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __global__ void lat_kernel(double *out, long long *cyc, double a, double b)
{
    double x = a;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) {
        if (OP == 0) x = fma(x, b, a);
        if (OP == 1) x = x + b;
        if (OP == 2) x = x * b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP> __global__ void thr_kernel(double *out, long long *cyc, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = a + j + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP> void thr(double *out, long long *cyc, int warps)
{
    const int iters = 2048;
    thr_kernel<ILP><<<148, warps * 32>>>(out, cyc, 1.0, 0.999999, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * ILP * warps;          // warp-level DFMAs per SM
    std::printf("  warps/SM %2d  ILP %d : %.3f warp-DFMA/cycle/SM (pipe peak 2.0)  -> %.0f%%\n", warps, ILP, ops / c, 50.0 * ops / c);
}

int main()
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const char *names[3] = { "DFMA", "DADD", "DMUL" };
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) lat_kernel<0><<<1, 32>>>(out, cyc, 1.0, 0.999999);
            if (op == 1) lat_kernel<1><<<1, 32>>>(out, cyc, 1.0, 0.999999);
            if (op == 2) lat_kernel<2><<<1, 32>>>(out, cyc, 1.0, 0.999999);
            cudaDeviceSynchronize();
        }
        long long c;
        cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
        std::printf("%s dependent-issue latency: %.2f cycles\n", names[op], c / 4096.0);
    }
    for (int warps : { 1, 2, 4, 7, 8, 12, 16 }) {
        thr<1>(out, cyc, warps);
        thr<2>(out, cyc, warps);
        thr<4>(out, cyc, warps);
        thr<8>(out, cyc, warps);
    }
    std::printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
