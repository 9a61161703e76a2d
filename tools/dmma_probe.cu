// dmma_probe.cu -- microbenchmarks that size the FP64 tensor (DMMA.8x8x4) pipe on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_probe tools/dmma_probe.cu
// (1) register-only DMMA loop: warps per SM x independent accumulators  -> issue rate per SMSP
// (2) LDS.64 fragment loads + DMMA (the inner loop of ee_gemm.cu without any global traffic)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NACC>
__global__ void reg_loop(double *out, int iters)
{
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp tile (8*MF) x (8*NF), fragments from shared memory with the ee_gemm layouts
template <int MF, int NF>
__global__ void lds_loop(double *out, int iters)
{
    extern __shared__ double sm[];
    constexpr int KC = 16, PA = 8 * MF * 4 + 4, PB = 8 * NF * 4 + 4;   // pretend 4 warps share along each dim
    for (int i = threadIdx.x; i < KC * (PA + PB); i += blockDim.x) sm[i] = i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const double *sa = sm + (warp & 3) * 8 * MF, *sb = sm + KC * PA + ((warp >> 2) & 3) * 8 * NF;
    double acc[NF][MF][2];
#pragma unroll
    for (int a = 0; a < NF; a++)
#pragma unroll
        for (int b = 0; b < MF; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double af[MF], bf[NF];
#pragma unroll
            for (int b = 0; b < MF; b++) af[b] = sa[(kk + fk) * PA + b * 8 + fi];
#pragma unroll
            for (int a = 0; a < NF; a++) bf[a] = sb[(kk + fk) * PB + a * 8 + fi];
#pragma unroll
            for (int a = 0; a < NF; a++)
#pragma unroll
                for (int b = 0; b < MF; b++) dmma(acc[a][b][0], acc[a][b][1], bf[a], af[b]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < NF; a++)
#pragma unroll
        for (int b = 0; b < MF; b++) s += acc[a][b][0] + acc[a][b][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    int nsm = 148;
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); nsm = pr.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * nsm * 1024 * 4);
    const int iters = 4000;
    printf("{\"sms\": %d,\n \"reg_loop\": [", nsm);
    bool first = true;
    auto rl = [&](auto kern, int nacc, int warps) {
        double ms = time_ms([&]() { kern<<<nsm, warps * 32>>>(out, iters); });
        double tf = (double)nsm * warps * nacc * iters * 512.0 / (ms * 1e-3) / 1e12;
        printf("%s\n  {\"warps_per_sm\": %d, \"nacc\": %d, \"tflops\": %.2f}", first ? "" : ",", warps, nacc, tf);
        first = false;
    };
    for (int w : {4, 8, 16, 32}) {
        rl(reg_loop<4>, 4, w); rl(reg_loop<8>, 8, w); rl(reg_loop<16>, 16, w); rl(reg_loop<32>, 32, w);
    }
    printf("],\n \"lds_loop\": [");
    first = true;
    auto ll = [&](auto kern, int mf, int nf, int warps) {
        size_t smem = 16 * (8 * mf * 4 + 4 + 8 * nf * 4 + 4) * sizeof(double);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        double ms = time_ms([&]() { kern<<<nsm, warps * 32, smem>>>(out, iters / 4); });
        double tf = (double)nsm * warps * mf * nf * 4 * (iters / 4) * 512.0 / (ms * 1e-3) / 1e12;
        printf("%s\n  {\"warps_per_sm\": %d, \"mf\": %d, \"nf\": %d, \"tflops\": %.2f}", first ? "" : ",", warps, mf, nf, tf);
        first = false;
    };
    for (int w : {4, 8, 16}) {
        ll(lds_loop<8, 4>, 8, 4, w); ll(lds_loop<4, 4>, 4, 4, w); ll(lds_loop<4, 2>, 4, 2, w); ll(lds_loop<2, 2>, 2, 2, w);
    }
    ll(lds_loop<4, 2>, 4, 2, 32); ll(lds_loop<2, 2>, 2, 2, 32);
    printf("]}\n");
    return 0;
}
