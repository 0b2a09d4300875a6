// ubench_pipes.cu -- issue-rate micro-benchmark for the add-compare-select instruction mix
// of the turbo-decoder kernels on sm_100a (B200).  Not part of the product; its numbers fix
// the ALU roofline denominator in DESIGN.md (SURVEY.md 8d asks to re-verify the B300 pipe
// table on B200).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
//   ./ubench_pipes            # prints one line per op: thread-ops / clk / SM
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHAINS 8
#define ITERS 512

enum Op {
    OP_FADD, OP_FMNMX, OP_FFMA, OP_FMNMX3, OP_IADD, OP_IMNMX, OP_IMAD, OP_LOP3, OP_PRMT, OP_SHF,
    OP_VIADD16, OP_VIMNMX16, OP_VIMNMX3_16, OP_VIADDMNMX16, OP_VIADDMNMX32, OP_HADD2, OP_HMNMX2, OP_HFMA2,
    OP_MIX_VIADDMNMX16_HADD2, OP_MIX_VIADDMNMX16_IMAD, OP_MIX_VIADD16_VIMNMX16, OP_MIX_FADD_FMNMX,
    OP_MIX_VIADDMNMX16_VIADD16, OP_MIX_VIMNMX16_IMAD, OP_MIX_VIADDMNMX16_LOP3, OP_MIX_HADD2_HMNMX2,
    OP_IMADHI, OP_IMADWIDE, OP_MIX_VIADDMNMX16_IMADHI, OP_MIX_VIADDMNMX16_IADD3, OP_MIX_VIADD16_IMAD, OP_MIX_VIADD16_HADD2, OP_MIX_VIADD16_FADD,
    OP_COUNT
};
static const char *op_name[OP_COUNT] = {
    "FADD", "FMNMX", "FFMA", "FMNMX3", "IADD3", "IMNMX(s32)", "IMAD", "LOP3", "PRMT", "SHF",
    "VIADD.16x2", "VIMNMX.S16x2", "VIMNMX3.S16x2", "VIADDMNMX.S16x2", "VIADDMNMX(s32)", "HADD2", "HMNMX2", "HFMA2",
    "mix VIADDMNMX.S16x2+HADD2", "mix VIADDMNMX.S16x2+IMAD", "mix VIADD.16x2+VIMNMX.S16x2", "mix FADD+FMNMX",
    "mix VIADDMNMX.S16x2+VIADD.16x2", "mix VIMNMX.S16x2+IMAD", "mix VIADDMNMX.S16x2+LOP3", "mix HADD2+HMNMX2",
    "IMAD.HI.U32", "IMAD.WIDE.U32", "mix VIADDMNMX.S16x2+IMAD.HI", "mix VIADDMNMX.S16x2+IADD3", "mix VIADD.16x2+IMAD", "mix VIADD.16x2+HADD2", "mix VIADD.16x2+FADD"};
// thread-ops counted per step() call
static const int op_per_step[OP_COUNT] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                          2, 2, 2, 2, 2, 2, 2, 2,
                                          1, 1, 2, 2, 2, 2, 2};

__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2 *>(&u); }

template <int OP>
__device__ __forceinline__ void step(unsigned &x, unsigned &w, unsigned y, unsigned z)
{
    // x, w are loop-carried; y, z are runtime constants.  Every op depends on x (or w).
    if (OP == OP_FADD) x = __float_as_uint(__uint_as_float(x) + __uint_as_float(y));
    if (OP == OP_FMNMX) { x = __float_as_uint(fminf(fmaxf(__uint_as_float(x), __uint_as_float(y)), __uint_as_float(z))); }  // two ops
    if (OP == OP_FFMA) x = __float_as_uint(fmaf(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z)));
    if (OP == OP_FMNMX3) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(__uint_as_float(x)), "f"(__uint_as_float(y)), "f"(__uint_as_float(z))); x = __float_as_uint(r); }
    if (OP == OP_IADD) { asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(z)); }  // ptxas fuses the pair into one IADD3
    if (OP == OP_IMNMX) x = (unsigned)min(max((int)x, (int)y), (int)z);  // two ops
    if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_VIADD16) x = __vadd2(x, y);
    if (OP == OP_VIMNMX16) { x = __vmaxs2(x, y); x = __vmins2(x, z); }  // two ops; counted below
    if (OP == OP_VIMNMX3_16) { x = __vimax3_s16x2(x, y, z); x = __vimin3_s16x2(x, z, y); }
    if (OP == OP_VIADDMNMX16) x = __viaddmax_s16x2(x, y, z);
    if (OP == OP_VIADDMNMX32) x = (unsigned)__viaddmax_s32((int)x, (int)y, (int)z);
    if (OP == OP_HADD2) x = h2u(__hadd2(u2h(x), u2h(y)));
    if (OP == OP_HMNMX2) { x = h2u(__hmax2(u2h(x), u2h(y))); x = h2u(__hmin2(u2h(x), u2h(z))); }
    if (OP == OP_HFMA2) x = h2u(__hfma2(u2h(x), u2h(y), u2h(z)));
    if (OP == OP_MIX_VIADDMNMX16_HADD2) { x = __viaddmax_s16x2(x, y, z); w = h2u(__hadd2(u2h(w), u2h(y))); }
    if (OP == OP_MIX_VIADDMNMX16_IMAD) { x = __viaddmax_s16x2(x, y, z); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w) : "r"(y), "r"(z)); }
    if (OP == OP_MIX_VIADD16_VIMNMX16) { x = __vadd2(x, y); w = __vmaxs2(w, x); }
    if (OP == OP_MIX_FADD_FMNMX) { x = __float_as_uint(__uint_as_float(x) + __uint_as_float(y)); w = __float_as_uint(fmaxf(__uint_as_float(w), __uint_as_float(x))); }
    if (OP == OP_MIX_VIADDMNMX16_VIADD16) { x = __viaddmax_s16x2(x, y, z); w = __vadd2(w, y); }
    if (OP == OP_MIX_VIMNMX16_IMAD) { x = __vmaxs2(x, y); x = __vmins2(x, z); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w) : "r"(y), "r"(z)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w) : "r"(z), "r"(y)); }
    if (OP == OP_MIX_VIADDMNMX16_LOP3) { x = __viaddmax_s16x2(x, y, z); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w) : "r"(y), "r"(z)); }
    if (OP == OP_IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_IMADWIDE) { unsigned long long r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(x), "r"(y), "l"((unsigned long long)z << 30)); x = (unsigned)(r >> 32); }
    if (OP == OP_MIX_VIADDMNMX16_IMADHI) { x = __viaddmax_s16x2(x, y, z); asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(w) : "r"(y), "r"(z)); }
    if (OP == OP_MIX_VIADDMNMX16_IADD3) { x = __viaddmax_s16x2(x, y, z); asm volatile("add.s32 %0, %0, %1;" : "+r"(w) : "r"(y)); asm volatile("add.s32 %0, %0, %1;" : "+r"(w) : "r"(z)); }
    if (OP == OP_MIX_VIADD16_IMAD) { x = __vadd2(x, y); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w) : "r"(y), "r"(z)); }
    if (OP == OP_MIX_VIADD16_HADD2) { x = __vadd2(x, y); w = h2u(__hadd2(u2h(w), u2h(y))); }
    if (OP == OP_MIX_VIADD16_FADD) { x = __vadd2(x, y); w = __float_as_uint(__uint_as_float(w) + __uint_as_float(y)); }
    if (OP == OP_MIX_HADD2_HMNMX2) { x = h2u(__hadd2(u2h(x), u2h(y))); w = h2u(__hmax2(u2h(w), u2h(x))); }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(const unsigned *in, unsigned *out, long long *cycles)
{
    unsigned x[CHAINS], w[CHAINS];
    unsigned y = in[0], z = in[1];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) { x[k] = in[2 + k] + threadIdx.x; w[k] = in[10 + k] ^ threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < CHAINS; k++) step<OP>(x[k], w[k], y, z);
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc ^= x[k] ^ w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const unsigned *d_in, unsigned *d_out, long long *d_cyc, int nsm, int threads)
{
    bench<OP><<<nsm, threads>>>(d_in, d_out, d_cyc);  // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<nsm, threads>>>(d_in, d_out, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(nsm);
    cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
    int per_step = op_per_step[OP];
    if (OP == OP_VIMNMX16 || OP == OP_VIMNMX3_16 || OP == OP_HMNMX2 || OP == OP_FMNMX || OP == OP_IMNMX) per_step = 2;
    if (OP == OP_MIX_VIMNMX16_IMAD) per_step = 4;
    double ops = (double)threads * CHAINS * 4.0 * ITERS * per_step;
    printf("%-34s threads/SM=%4d  %7.2f thread-ops/clk/SM   (%.3f ms, %.0f MHz eff)\n", op_name[OP], threads,
           ops / (double)mx, ms, (double)mx / (ms * 1e3));
}

template <int OP>
struct Runner {
    static void go(const unsigned *a, unsigned *b, long long *c, int nsm, int th, int only)
    {
        if (only < 0 || only == OP) run<OP>(a, b, c, nsm, th);
        Runner<OP + 1>::go(a, b, c, nsm, th, only);
    }
};
template <>
struct Runner<OP_COUNT> {
    static void go(const unsigned *, unsigned *, long long *, int, int, int) {}
};

int main(int argc, char **argv)
{
    int only = argc > 1 ? atoi(argv[1]) : -1;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nsm = prop.multiProcessorCount;
    printf("%s  SMs=%d  clock=%d kHz\n", prop.name, nsm, prop.clockRate);
    unsigned h_in[32];
    for (int i = 0; i < 32; i++) h_in[i] = 0x00030001u * (i + 1);
    h_in[0] = 0x3c003c00u;  // y: 1.0h,1.0h / small ints
    h_in[1] = 0x00050003u;
    unsigned *d_in, *d_out; long long *d_cyc;
    cudaMalloc(&d_in, sizeof(h_in)); cudaMalloc(&d_out, sizeof(unsigned) * nsm * 1024); cudaMalloc(&d_cyc, sizeof(long long) * nsm);
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    for (int th : {1024, 256}) Runner<0>::go(d_in, d_out, d_cyc, nsm, th, only);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
