// tdb200_ratematch.cu -- TS 36.212 rate matching around the decode path: the stage the
// reference declares as rate_match() / de_rate_match() (ITTC/main.h:23-24; call sites commented out
// at ITTC/main.cpp:196,204) and never wrote.
//
// The whole of 36.212 5.1.3.2.2 (tail-bit multiplexing into d0, d1, d2) and 5.1.4.1 (sub-block
// interleavers, circular buffer, bit selection from k0(rv) with <NULL> pruning) collapses into one
// permutation per (K, rv, Ncb): `perm[j]` = position in the reference's multiplex order
// (log_map.cpp:566-578) of the j-th transmitted bit, j < nnn = number of non-<NULL> buffer entries,
// after which the sequence repeats.  build_rm_table() derives it in closed form (no padded matrix);
// the oracle (oracle/turbo_oracle_rm.c) restates the specification's matrices and loop literally.
//
//   rate_match_kernel     e[cb][k] = coded[cb][perm[k mod nnn]]                         (bits, gather)
//   rate_dematch_kernel   llr[cb][n] = (old +) sum_{k = inv[n] + m*nnn < E} e[cb][k]      (soft values)
//
// De-rate-matching is a gather through an irregular permutation: one CTA per codeblock stages one wrap
// of the received row in shared memory with coalesced loads and gathers from there, so HBM sees each
// byte once and every global store is coalesced.  Repeated bits (E > nnn) are combined in
// transmission order; punctured ones come out as 0 (no information); `accumulate` adds to what the
// output already holds (HARQ combining of retransmissions with other rv): new = old + (this
// transmission's sum).  Accumulation is fp32
// (fp64 for double input, saturating integers for 8-bit input).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "tdb200_internal.h"

namespace tdb200 {

// ---------------------------------------------------------------- host: the permutation
bool build_rm_table(int K, int rv, int Ncb, int F, std::vector<int> &perm, std::vector<int> &inv)
{
    static const int P[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                              1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
    const int D = K + 4, R = (D + 31) / 32, Kpi = 32 * R, ND = Kpi - D, Kw = 3 * Kpi, NL = 3 * K + 12;
    if (Ncb <= 0 || Ncb > Kw) Ncb = Kw;
    // element k of stream s -> multiplex position (tail order of 36.212 5.1.3.2.2:
    // d0 = x_K z_K+1 x'_K z'_K+1, d1 = z_K x_K+2 z'_K x'_K+2, d2 = x_K+1 z_K+2 x'_K+1 z'_K+2)
    auto mux = [&](int s, int k) -> int {
        if (k < K) return 3 * k + s;
        // offsets from 3K in the reference's order: x_K+m at 2m, z_K+m at 2m+1, x'_K+m at 6+2m, z'_K+m at 7+2m
        static const int off[3][4] = {{0, 3, 6, 9}, {1, 4, 7, 10}, {2, 5, 8, 11}};
        return 3 * K + off[s][k - K];
    };
    auto source = [&](int p) -> int {  // circular-buffer position -> multiplex position, -1 for <NULL>
        int s, k, y;
        if (p < Kpi) { s = 0; k = p; }
        else { s = 1 + ((p - Kpi) & 1); k = (p - Kpi) >> 1; }
        if (s < 2) y = (k % R) * 32 + P[k / R];
        else y = (P[k / R] + 32 * (k % R) + 1) % Kpi;
        // <NULL>: the interleaver's dummy bits, and d0 / d1 of the F filler bits (36.212 5.1.3.2.1)
        return (y < ND || (s < 2 && y - ND < F)) ? -1 : mux(s, y - ND);
    };
    const long k0 = (long)R * (2L * ((Ncb + 8 * R - 1) / (8 * R)) * rv + 2);
    perm.clear();
    inv.assign(NL, -1);
    for (int k = 0; k < F; k++) inv[3 * k] = inv[3 * k + 1] = kRmFiller;  // known zeros: the soft inverse writes a fixed value there
    for (int j = 0; j < Ncb; j++) {
        const int n = source((int)((k0 + j) % Ncb));
        if (n >= 0) { inv[n] = (int)perm.size(); perm.push_back(n); }
    }
    return !perm.empty();
}

namespace {

template <int T> struct ElemOf;
template <> struct ElemOf<TDB200_LLR_F64> { using type = double; };
template <> struct ElemOf<TDB200_LLR_F32> { using type = float; };
template <> struct ElemOf<TDB200_LLR_F16> { using type = __half; };
template <> struct ElemOf<TDB200_LLR_S8> { using type = int8_t; };

template <typename A, int T> __device__ __forceinline__ A load_as(const void *p, size_t i)
{
    using E = typename ElemOf<T>::type;
    const E v = static_cast<const E *>(p)[i];
    if constexpr (T == TDB200_LLR_F16) return (A)__half2float(v);
    else return (A)v;
}

// clamp(rint(x * scale), +-clip), NaN -> 0, for clip <= 127: cvt.rni.sat.s8 rounds to nearest even, saturates to
// [-128, 127] and maps NaN to 0 -- the decoder's quantiser (quant() in tdb200_fast_kernel.cuh) in two fewer steps
__device__ __forceinline__ int quant8(float x, float scale, int clip)
{
    int q;
    asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(x * scale));
    return max(min(q, clip), -clip);
}

// Shared-memory word index of logical word j: the sub-block interleaver sends the 32 lanes of a gather to
// addresses that differ by multiples of R (eight-way bank conflicts on average at K = 6144); XOR-ing the
// row number into the bank spreads them (2.2-way), and a linear fill stays conflict-free.
__device__ __forceinline__ int swz(int j) { return j ^ ((j >> 5) & 31); }

// acc (in the accumulation type of IN_T) -> element of OUT_T.  IN_T == OUT_T is the plain inverse;
// OUT_T == S8 from a float type applies the throughput decoder's channel quantiser to exactly the
// value the plain inverse would have stored (so the two-call form gives the same decode).
template <int IN_T, int OUT_T, typename A>
__device__ __forceinline__ void store_as(void *p, size_t i, A acc, float scale, int clip)
{
    if constexpr (OUT_T == TDB200_LLR_F64) static_cast<double *>(p)[i] = (double)acc;
    else if constexpr (OUT_T == TDB200_LLR_F32) static_cast<float *>(p)[i] = (float)acc;
    else if constexpr (OUT_T == TDB200_LLR_F16) static_cast<__half *>(p)[i] = __float2half_rn((float)acc);
    else if constexpr (IN_T == TDB200_LLR_S8) static_cast<int8_t *>(p)[i] = (int8_t)max(min((int)acc, 127), -127);
    else if constexpr (IN_T == TDB200_LLR_F16) static_cast<int8_t *>(p)[i] = (int8_t)quant8(__half2float(__float2half_rn((float)acc)), scale, clip);
    else static_cast<int8_t *>(p)[i] = (int8_t)quant8((float)acc, scale, clip);
}

__device__ __forceinline__ unsigned char coded_at(const unsigned char *sm, int n) { return sm[4 * swz(n >> 2) + (n & 3)]; }

// One CTA per codeblock: the coded row is staged in shared memory (coalesced 16-byte loads when the
// row is aligned), the transmitted bits are gathered from there and leave four per store.
__global__ void __launch_bounds__(256) rate_match_kernel(const uint8_t *__restrict__ coded, uint8_t *__restrict__ e_bits, const int *__restrict__ perm,
                                                         int nnn, int NL, int E)
{
    extern __shared__ __align__(16) unsigned char rm_smem[];
    const size_t cb = blockIdx.x;
    const uint8_t *row = coded + cb * (size_t)NL;
    unsigned *stage = reinterpret_cast<unsigned *>(rm_smem);  // words swizzled like the soft inverse's tile
    if ((reinterpret_cast<size_t>(row) & 3) == 0) {  // NL is a multiple of 4
        const unsigned *src = reinterpret_cast<const unsigned *>(row);
#pragma unroll 8
        for (int i = threadIdx.x; i < NL / 4; i += blockDim.x) stage[swz(i)] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < NL; i += blockDim.x) rm_smem[4 * swz(i >> 2) + (i & 3)] = row[i];
    }
    __syncthreads();
    uint8_t *out = e_bits + cb * (size_t)E;
    if ((reinterpret_cast<size_t>(out) & 3) == 0) {
#pragma unroll 4
        for (int k4 = threadIdx.x; k4 < E / 4; k4 += blockDim.x) {
            unsigned v = 0;
            int j = (4 * k4) % nnn;  // one division per four bits; the wrap is a compare
#pragma unroll
            for (int b = 0; b < 4; b++) {
                v |= (unsigned)coded_at(rm_smem, __ldg(perm + j)) << (8 * b);
                j = (j + 1 == nnn) ? 0 : j + 1;
            }
            reinterpret_cast<unsigned *>(out)[k4] = v;
        }
        for (int k = (E & ~3) + threadIdx.x; k < E; k += blockDim.x) out[k] = coded_at(rm_smem, __ldg(perm + k % nnn));
    } else {
        for (int k = threadIdx.x; k < E; k += blockDim.x) out[k] = coded_at(rm_smem, __ldg(perm + k % nnn));
    }
}

constexpr int kRmThreads = 512;

// One CTA per codeblock.  `tile[j]` collects the soft value of transmitted position j: the first wrap
// of the circular buffer is loaded, later wraps (repetition) are added in place -- transmission order,
// coalesced loads.  Then every multiplex position n gathers tile[inv[n]] (0 if never sent), four
// consecutive n per thread so that the stores are 4 to 32 bytes wide.
template <int IN_T, int OUT_T>
__global__ void __launch_bounds__(kRmThreads) rate_dematch_kernel(const void *__restrict__ e_llr, void *llr, const int *__restrict__ inv, int nnn, int NL,
                                                                  int E, int accumulate, float scale, int clip)
{
    using A = typename std::conditional<IN_T == TDB200_LLR_F64, double, typename std::conditional<IN_T == TDB200_LLR_S8, int, float>::type>::type;
    extern __shared__ __align__(16) unsigned char rm_smem[];
    A *tile = reinterpret_cast<A *>(rm_smem);
    const size_t cb = blockIdx.x;
    const size_t in0 = cb * (size_t)E, out0 = cb * (size_t)NL;
    const int len = min(nnn, E);
    // both phases are latency-bound streams: keep eight (four) independent loads in flight per thread
#pragma unroll 8
    for (int k = threadIdx.x; k < len; k += kRmThreads) tile[swz(k)] = load_as<A, IN_T>(e_llr, in0 + k);
    for (int w0 = nnn; w0 < E; w0 += nnn) {  // repetition: position k stays with the thread that loaded it
        const int wl = min(nnn, E - w0);
#pragma unroll 8
        for (int k = threadIdx.x; k < wl; k += kRmThreads) tile[swz(k)] += load_as<A, IN_T>(e_llr, in0 + w0 + k);
    }
    __syncthreads();
#pragma unroll 4
    for (int n4 = threadIdx.x; n4 < NL / 4; n4 += kRmThreads) {  // NL is a multiple of 4
        const int4 j4 = __ldg(reinterpret_cast<const int4 *>(inv) + n4);
        const int js[4] = {j4.x, j4.y, j4.z, j4.w};
        A v[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            v[b] = (js[b] >= 0 && js[b] < len) ? tile[swz(js[b])] : (A)0;
            if (accumulate) v[b] = load_as<A, OUT_T>(llr, out0 + 4 * n4 + b) + v[b];
            if (js[b] == kRmFiller) v[b] = (A)(IN_T == TDB200_LLR_S8 ? -127 : kRmFillerLlr);  // a filler bit (or its parity-1 bit): known 0
        }
        if constexpr (OUT_T == TDB200_LLR_S8) {
            unsigned w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int q;
                if constexpr (IN_T == TDB200_LLR_S8) q = max(min((int)v[b], 127), -127);
                else if constexpr (IN_T == TDB200_LLR_F16) q = quant8(__half2float(__float2half_rn((float)v[b])), scale, clip);
                else q = quant8((float)v[b], scale, clip);
                w |= ((unsigned)q & 0xffu) << (8 * b);
            }
            reinterpret_cast<unsigned *>(static_cast<int8_t *>(llr) + out0)[n4] = w;
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++) store_as<IN_T, OUT_T, A>(llr, out0 + 4 * n4 + b, v[b], scale, clip);
        }
    }
}

template <int IN_T, int OUT_T>
cudaError_t dematch_launch(const RmArgs &a, cudaStream_t st)
{
    using A = typename std::conditional<IN_T == TDB200_LLR_F64, double, typename std::conditional<IN_T == TDB200_LLR_S8, int, float>::type>::type;
    const size_t smem = sizeof(A) * (size_t)((std::min(a.nnn, std::max(a.E, 1)) + 31) & ~31);  // swz() permutes within rows of 32 words
    // the attribute belongs to the kernel, not to a call: always opt in to the device maximum
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e == cudaSuccess && smem > (size_t)optin) e = cudaErrorInvalidValue;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rate_dematch_kernel<IN_T, OUT_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e != cudaSuccess) return e;
    rate_dematch_kernel<IN_T, OUT_T><<<a.n_cb, kRmThreads, smem, st>>>(a.e_llr, a.llr, a.inv, a.nnn, a.NL, a.E, a.accumulate,
                                                                      (float)(1 << a.frac_bits), a.clip);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_rate_match(const uint8_t *coded, uint8_t *e_bits, const int *perm, int nnn, int NL, int E, int n_cb, cudaStream_t st)
{
    if (n_cb == 0 || E == 0) return cudaSuccess;
    rate_match_kernel<<<n_cb, 256, (NL + 127) & ~127, st>>>(coded, e_bits, perm, nnn, NL, E);
    return cudaGetLastError();
}

cudaError_t launch_rate_dematch(const RmArgs &a, cudaStream_t st)
{
    if (a.n_cb == 0) return cudaSuccess;
    if (a.NL % 4) return cudaErrorInvalidValue;
    const int in = a.in_type, out = a.out_type;
    if (in == out) {
        switch (in) {
            case TDB200_LLR_F64: return dematch_launch<TDB200_LLR_F64, TDB200_LLR_F64>(a, st);
            case TDB200_LLR_F32: return dematch_launch<TDB200_LLR_F32, TDB200_LLR_F32>(a, st);
            case TDB200_LLR_F16: return dematch_launch<TDB200_LLR_F16, TDB200_LLR_F16>(a, st);
            default: return dematch_launch<TDB200_LLR_S8, TDB200_LLR_S8>(a, st);
        }
    }
    if (out != TDB200_LLR_S8) return cudaErrorInvalidValue;  // the only cross-type hand-over: to the s16 decoder
    switch (in) {
        case TDB200_LLR_F64: return dematch_launch<TDB200_LLR_F64, TDB200_LLR_S8>(a, st);
        case TDB200_LLR_F32: return dematch_launch<TDB200_LLR_F32, TDB200_LLR_S8>(a, st);
        default: return dematch_launch<TDB200_LLR_F16, TDB200_LLR_S8>(a, st);
    }
}

}  // namespace tdb200
