// ubench_fp64.cu -- latency and issue rate of the fp64 instructions the reference-order decoder
// (csrc/tdb200_ref64.cu) is made of, on sm_100a (B200).  Not part of the product; its numbers
// decide how that kernel is laid out.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu && ./ubench_fp64
//
// latency:    one warp, one dependent chain            -> clk per operation
// throughput: 1024 threads per SM, 8 chains per thread -> thread-ops / clk / SM
#include <cuda_runtime.h>
#include <cstdio>

enum Op { DADD, DMAX, DSETP_SEL, DADD_ABS, SHFL64, LDS_RT, ISETP64_SEL, OPS };
static const char *names[OPS] = {"DADD", "max.f64 (DSETP+SEL or DMNMX)", "DSETP + 2xSEL (ternary on doubles)", "|x-y| (DADD + abs)",
                                 "64-bit shuffle (width 8)", "STS.64 + LDS.64 round trip", "64-bit integer compare + select"};

template <int OP>
__device__ __forceinline__ double step(double x, double y, double z, double *sm)
{
    if (OP == DADD) return x + y;
    if (OP == DMAX) return fmax(x, y) + 0.0 * z;  // keep z alive without a second dependent op (folded away)
    if (OP == DSETP_SEL) return x < y ? z : x;
    if (OP == DADD_ABS) return fabs(x - y);
    if (OP == SHFL64) return __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 7, 8);
    if (OP == LDS_RT) {
        sm[threadIdx.x] = x;
        __syncwarp();
        double r = sm[threadIdx.x ^ 1];
        __syncwarp();
        return r;
    }
    if (OP == ISETP64_SEL) {
        long long a = __double_as_longlong(x), b = __double_as_longlong(y);
        return __longlong_as_double(a < b ? __double_as_longlong(z) : a);
    }
    return x;
}

template <int OP, int CHAINS>
__global__ void __launch_bounds__(1024, 1) bench(const double *in, double *out, long long *cycles, int iters)
{
    __shared__ double sm[1024];
    double v[CHAINS];
    const double y = in[1], z = in[2];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) v[c] = in[0] + c + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) v[c] = step<OP>(v[c], y, z, sm);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const double *din, double *dout, long long *dcyc, int sms)
{
    const int iters = 256;
    long long c;
    bench<OP, 1><<<1, 32>>>(din, dout, dcyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, dcyc, sizeof(c), cudaMemcpyDeviceToHost);
    double lat = (double)c / (iters * 8);
    bench<OP, 8><<<sms, 1024>>>(din, dout, dcyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, dcyc, sizeof(c), cudaMemcpyDeviceToHost);
    double rate = 1024.0 * 8 * iters * 8 / (double)c;
    std::printf("%-40s latency %6.1f clk   rate %6.1f thread-ops/clk/SM\n", names[OP], lat, rate);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    std::printf("%s  SMs=%d\n", p.name, p.multiProcessorCount);
    double h[3] = {1.0, 1e-9, 0.5}, *din, *dout;
    long long *dcyc;
    cudaMalloc(&din, sizeof(h));
    cudaMalloc(&dout, sizeof(double) * 1024 * p.multiProcessorCount);
    cudaMalloc(&dcyc, sizeof(long long) * p.multiProcessorCount);
    cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    run<DADD>(din, dout, dcyc, p.multiProcessorCount);
    run<DMAX>(din, dout, dcyc, p.multiProcessorCount);
    run<DSETP_SEL>(din, dout, dcyc, p.multiProcessorCount);
    run<DADD_ABS>(din, dout, dcyc, p.multiProcessorCount);
    run<SHFL64>(din, dout, dcyc, p.multiProcessorCount);
    run<LDS_RT>(din, dout, dcyc, p.multiProcessorCount);
    run<ISETP64_SEL>(din, dout, dcyc, p.multiProcessorCount);
    return cudaDeviceSynchronize() != cudaSuccess;
}
