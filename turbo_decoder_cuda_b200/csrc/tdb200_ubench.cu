// tdb200_ubench.cu -- issue-rate micro-benchmark of the add-compare-select instruction mix (the measured
// denominator of the ALU roofline that bench.py reports; tools/ubench_pipes.cu is the long form with every pipe).
//
// One CTA of 1024 threads per SM runs eight independent dependency chains per thread of
//     x = max(x + y, z)   VIADDMNMX.S16x2  (ALU pipe)        w = w + y   VIADD.16x2  (fma-heavy pipe)
// -- the two instructions the recursions of fast_s16_kernel are made of, on the two pipes they issue to -- and
// reports thread-operations per clock per SM from clock64() around the loop.  128 would be one warp-instruction
// per clock on each of the four sub-partitions; B200 delivers about 120 for this mix and 64 for either alone.
#include <cuda_runtime.h>

#include <vector>

#include "tdb200.h"

namespace {

constexpr int kChains = 8, kIters = 512, kUnroll = 4;

template <int MIX>
__global__ void __launch_bounds__(1024, 1) issue_rate_kernel(const unsigned *in, unsigned *out, long long *cycles)
{
    unsigned x[kChains], w[kChains];
    const unsigned y = in[0], z = in[1];
#pragma unroll
    for (int k = 0; k < kChains; k++) { x[k] = in[2 + k] + threadIdx.x; w[k] = in[10 + k] ^ threadIdx.x; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
#pragma unroll
            for (int k = 0; k < kChains; k++) {
                if (MIX != 1) x[k] = __viaddmax_s16x2(x[k], y, z);
                if (MIX != 0) w[k] = __vadd2(w[k], y);
            }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < kChains; k++) acc ^= x[k] ^ w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace

extern "C" int tdb200_ubench_issue_rate(int device, int mix, double *thread_ops_per_clk_per_sm)
{
    if (!thread_ops_per_clk_per_sm || mix < 0 || mix > 2) return TDB200_ERR_INVALID_ARG;
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return TDB200_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    const int nsm = prop.multiProcessorCount;
    unsigned h_in[32];
    for (int i = 0; i < 32; i++) h_in[i] = 0x00030001u * (i + 1);
    unsigned *d_in = nullptr, *d_out = nullptr;
    long long *d_cyc = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&d_in, sizeof(h_in));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(unsigned) * nsm * 1024);
    if (e == cudaSuccess) e = cudaMalloc(&d_cyc, sizeof(long long) * nsm);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    std::vector<long long> cyc(nsm, 0);
    for (int rep = 0; rep < 2 && e == cudaSuccess; rep++) {  // the first launch warms the instruction cache
        if (mix == 0) issue_rate_kernel<0><<<nsm, 1024>>>(d_in, d_out, d_cyc);
        else if (mix == 1) issue_rate_kernel<1><<<nsm, 1024>>>(d_in, d_out, d_cyc);
        else issue_rate_kernel<2><<<nsm, 1024>>>(d_in, d_out, d_cyc);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess) e = cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_cyc);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return TDB200_ERR_CUDA;
    long long mx = 1;
    for (long long c : cyc) mx = c > mx ? c : mx;
    const double ops = 1024.0 * kChains * kUnroll * kIters * (mix == 2 ? 2 : 1);
    *thread_ops_per_clk_per_sm = ops / (double)mx;
    return TDB200_OK;
}
