// tdb200_modem.cu -- the mapper and the soft demapper either side of the decode path,
// device-side, batched.  Compiled with -fmad=false: the demappers are specified operation by operation.
//
//   modulate_kernel   module()     ITTC/modanddem.cpp:175-186 (tables :7-71)   bit groups -> (I, Q)
//   awgn_kernel       AWGN()       ITTC/log_map.cpp:1388 (y = x + sigma * n); n is Philox4x32-10 +
//                                  Box-Muller, a pure function of (seed, element index), not the
//                                  reference's rand()-based central-limit sum
//   demap64_kernel    demodule()   ITTC/modanddem.cpp:674-686 in the reference's own order of operations
//                                  (per bit one ascending scan over all 2^M points, fp64): bit-identical
//                                  to the reference, feeds TDB200_ALGO_LOGMAP_F64
//   demap32_kernel    the same max-log metric LLR_b = -Kf (min_{bit b = 1} d - min_{bit b = 0} d) in fp32,
//                     per axis where the constellation is a product of two level sets (BPSK, QPSK,
//                     16QAM, 64QAM: the other axis' minimum cancels in the difference), exhaustive for
//                     8PSK; writes float, half, or the throughput decoder's 8-bit fixed-point channel
//                     values (clamp(rint(LLR * 2^F)), the same quantiser as the decoder's own load stage).
//                     oracle: tdo_demap_f32 / tdo_quant_s8 (oracle/turbo_oracle_mod.c), bit-exact.
//
// All four are flat element-wise kernels over [n_cb][3K+12] bits / LLRs and [n_cb][(3K+12)/M] symbols
// (rows are contiguous, and 3K+12 is a multiple of 12 for every K that is a multiple of 8, so a group
// of 12 LLRs never straddles a symbol or a row).  HBM-bound: a thread of demap32_kernel produces 12
// LLRs from 12/M symbols -- 16-byte stores, all sectors of the loads used by the warp.
#include <algorithm>

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

// Level sets (index = the axis' bits, MSB first) and the 8PSK points (index built LSB first, :136).
__constant__ double c_lv[4][8] = {
    {-1.0, 1.0},                                                            // BPSK  :7-10
    {0.7071, -0.7071},                                                      // QPSK  :17-24
    {-0.948683, -0.316228, 0.948683, 0.316228},                             // 16QAM :35-49
    {0.4629, 0.1543, 0.7615, 1.0801, -0.4629, -0.1543, -0.7615, -1.0801}};  // 64QAM :51-71
__constant__ double c_psk_i[8] = {-0.7071, -1, 0, 0.7071, 0, -0.7071, 0.7071, 1};   // :26-29
__constant__ double c_psk_q[8] = {0.7071, 0, 1, 0.7071, -1, -0.7071, -0.7071, 0};   // :31-34

__host__ __device__ constexpr int lv_row(int M) { return M == 1 ? 0 : (M == 2 ? 1 : (M == 4 ? 2 : 3)); }
__host__ __device__ constexpr int q_bits(int M) { return M == 1 ? 0 : (M == 3 ? 0 : M / 2); }  // bits on the Q axis

template <typename T> __device__ __forceinline__ float ld_f(const T *p, size_t i);
template <> __device__ __forceinline__ float ld_f<float>(const float *p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ld_f<double>(const double *p, size_t i) { return (float)__ldg(p + i); }
template <> __device__ __forceinline__ float ld_f<__half>(const __half *p, size_t i) { return __half2float(__ldg(p + i)); }
template <typename T> __device__ __forceinline__ double ld_d(const T *p, size_t i);
template <> __device__ __forceinline__ double ld_d<float>(const float *p, size_t i) { return (double)__ldg(p + i); }
template <> __device__ __forceinline__ double ld_d<double>(const double *p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ double ld_d<__half>(const __half *p, size_t i) { return (double)__half2float(__ldg(p + i)); }
template <typename T> __device__ __forceinline__ void st_f(T *p, size_t i, float v);
template <> __device__ __forceinline__ void st_f<float>(float *p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st_f<double>(double *p, size_t i, float v) { p[i] = (double)v; }
template <> __device__ __forceinline__ void st_f<__half>(__half *p, size_t i, float v) { p[i] = __float2half_rn(v); }

// ---------------------------------------------------------------- mapper
template <int M, typename T>
__global__ void __launch_bounds__(256) modulate_kernel(const uint8_t *__restrict__ coded, T *__restrict__ si, T *__restrict__ sq, size_t n_sym)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sym) return;
    int j = 0;
    if (M == 3) j = (coded[3 * s + 2] & 1) * 4 + (coded[3 * s + 1] & 1) * 2 + (coded[3 * s] & 1);
    else {
#pragma unroll
        for (int b = 0; b < M; b++) j = 2 * j + (coded[M * s + b] & 1);
    }
    double pi, pq;
    if (M == 3) { pi = c_psk_i[j]; pq = c_psk_q[j]; }
    else {
        constexpr int nq = q_bits(M);
        pi = c_lv[lv_row(M)][j >> nq];
        pq = nq ? c_lv[lv_row(M)][j & ((1 << nq) - 1)] : 0.0;
    }
    if constexpr (sizeof(T) == 8) {  // the reference's doubles, exactly
        si[s] = pi;
        sq[s] = pq;
    } else {
        st_f(si, s, (float)pi);
        st_f(sq, s, (float)pq);
    }
}

// ---------------------------------------------------------------- additive white Gaussian noise
template <typename T>
__global__ void __launch_bounds__(256) awgn_kernel(const T *__restrict__ x, T *__restrict__ y, size_t n, float sigma, unsigned long long seed)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // four elements per thread
    const size_t i0 = 4 * q;
    if (i0 >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, /*subsequence*/ q, /*offset*/ 0, &st);
    const float4 g = curand_normal4(&st);
    const float nz[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (i0 + j < n) st_f(y, i0 + j, ld_f(x, i0 + j) + sigma * nz[j]);
}

// ---------------------------------------------------------------- demodule() in reference order, fp64
template <int M, typename T>
__global__ void __launch_bounds__(256) demap64_kernel(const T *__restrict__ si, const T *__restrict__ sq, double *__restrict__ out, size_t n_sym, double kf)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sym) return;
    const double x = ld_d(si, s), y = ld_d(sq, s);
    double d[1 << M];
#pragma unroll
    for (int j = 0; j < (1 << M); j++) {
        double pi, pq;
        if (M == 3) { pi = c_psk_i[j]; pq = c_psk_q[j]; }
        else {
            constexpr int nq = q_bits(M);
            pi = c_lv[lv_row(M)][j >> nq];
            pq = nq ? c_lv[lv_row(M)][j & ((1 << nq) - 1)] : 0.0;
        }
        const double dr = x - pi, di = y - pq;
        d[j] = dr * dr + di * di;  // calculate_sqr_dis, :73-86 (no contraction: -fmad=false)
    }
#pragma unroll
    for (int b = 0; b < M; b++) {
        const int mask = (M == 3) ? (1 << b) : (1 << (M - 1 - b));
        double m1 = (M <= 2) ? (double)0x7fffffffffff : (double)0x7fffffff, m0 = m1;  // :198,237 / :304,394,510
#pragma unroll
        for (int j = 0; j < (1 << M); j++) {
            if (j & mask) { if (d[j] < m1) m1 = d[j]; }
            else          { if (d[j] < m0) m0 = d[j]; }
        }
        out[M * s + b] = -kf * (m1 - m0);
    }
}

// ---------------------------------------------------------------- fp32 demapper, 12 LLRs per thread
__device__ __forceinline__ float sqf(float a) { return a * a; }

template <int NB>
__device__ __forceinline__ void axis32(float v, const double *lv, float kf, float *out)
{
    float d[1 << NB];
#pragma unroll
    for (int l = 0; l < (1 << NB); l++) d[l] = sqf(v - (float)lv[l]);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int mask = (1 << NB) >> (b + 1);
        float m1 = 0.f, m0 = 0.f;
        bool h1 = false, h0 = false;
#pragma unroll
        for (int l = 0; l < (1 << NB); l++) {
            if (l & mask) { m1 = h1 ? fminf(m1, d[l]) : d[l]; h1 = true; }
            else          { m0 = h0 ? fminf(m0, d[l]) : d[l]; h0 = true; }
        }
        out[b] = -kf * (m1 - m0);
    }
}

// clamp(rint(x * scale), +-clip), NaN -> 0, for clip <= 127: cvt.rni.sat.s8 rounds to nearest even, saturates to
// [-128, 127] and maps NaN to 0 -- the decoder's quantiser (quant() in tdb200_fast_kernel.cuh) in two fewer steps
__device__ __forceinline__ int quant8(float x, float scale, int clip)
{
    int q;
    asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(x * scale));
    return max(min(q, clip), -clip);
}

// the M soft bits of one received symbol (x, y)
template <int M>
__device__ __forceinline__ void demap_symbol(float x, float y, float kf, float *llr)
{
    if (M == 3) {
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = sqf(x - (float)c_psk_i[j]) + sqf(y - (float)c_psk_q[j]);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            float m1 = 0.f, m0 = 0.f;
            bool h1 = false, h0 = false;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j & (1 << b)) { m1 = h1 ? fminf(m1, d[j]) : d[j]; h1 = true; }
                else              { m0 = h0 ? fminf(m0, d[j]) : d[j]; h0 = true; }
            }
            llr[b] = -kf * (m1 - m0);
        }
    } else {
        constexpr int nq = q_bits(M), ni = M - nq;
        axis32<ni>(x, c_lv[lv_row(M)], kf, llr);
        if (nq) axis32<(nq ? nq : 1)>(y, c_lv[lv_row(M)], kf, llr + ni);
    }
}

// the last n_llr % 12 soft bits of a flat call (rows of arbitrary length): one thread per symbol
template <int M, typename T>
__global__ void __launch_bounds__(256) demap32_tail_kernel(const T *__restrict__ si, const T *__restrict__ sq, void *__restrict__ out, int out_type,
                                                           size_t first_sym, size_t n_sym, float kf, float scale, int clip)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_sym) return;
    float llr[M];
    demap_symbol<M>(ld_f(si, first_sym + t), (M == 1) ? 0.f : ld_f(sq, first_sym + t), kf, llr);
#pragma unroll
    for (int b = 0; b < M; b++) {
        const size_t o = (first_sym + t) * M + b;
        if (out_type == TDB200_LLR_S8) static_cast<int8_t *>(out)[o] = (int8_t)quant8(llr[b], scale, clip);
        else if (out_type == TDB200_LLR_F16) static_cast<__half *>(out)[o] = __float2half_rn(llr[b]);
        else static_cast<float *>(out)[o] = llr[b];
    }
}

template <int M, typename T, int OUT_T>
__global__ void __launch_bounds__(256) demap32_kernel(const T *__restrict__ si, const T *__restrict__ sq, void *__restrict__ out, size_t n_groups,
                                                      float kf, float scale, int clip)
{
    __shared__ __align__(16) unsigned stage[OUT_T == TDB200_LLR_S8 ? 3 * 256 : 4];
    const size_t g0 = (size_t)blockIdx.x * blockDim.x;  // first group of this CTA
    const size_t gi = g0 + threadIdx.x;
    const bool live = gi < n_groups;
    constexpr int NS = 12 / M;
    float llr[12];
    float xs[NS], ys[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        xs[k] = live ? ld_f(si, gi * NS + k) : 0.f;
        ys[k] = (M == 1 || !live) ? 0.f : ld_f(sq, gi * NS + k);
    }
#pragma unroll
    for (int k = 0; k < NS; k++) demap_symbol<M>(xs[k], ys[k], kf, llr + M * k);
    if (OUT_T == TDB200_LLR_S8) {
        // 12 bytes per thread: staged through shared memory (word stride 3: conflict-free) so that the
        // CTA's 3072 contiguous output bytes leave as 16-byte stores
#pragma unroll
        for (int k = 0; k < 3; k++) {
            unsigned v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) v |= ((unsigned)quant8(llr[4 * k + j], scale, clip) & 0xffu) << (8 * j);
            stage[3 * threadIdx.x + k] = v;
        }
        __syncthreads();
        const size_t left = n_groups - g0;
        const int words = 3 * (int)(left < 256 ? left : 256);
        unsigned *o = reinterpret_cast<unsigned *>(static_cast<int8_t *>(out) + 12 * g0);  // 3072 * blockIdx: 16-byte aligned
        const int t = threadIdx.x;
        const bool al16 = (reinterpret_cast<size_t>(o) & 15) == 0;  // caller-supplied buffers may be row-offset views
        if (al16 && 4 * t + 3 < words) reinterpret_cast<uint4 *>(o)[t] = reinterpret_cast<const uint4 *>(stage)[t];
        else
            for (int w = 4 * t; w < min(words, 4 * t + 4); w++) o[w] = stage[w];
    } else if (OUT_T == TDB200_LLR_F16) {
        if (!live) return;
        uint2 *o = reinterpret_cast<uint2 *>(static_cast<__half *>(out) + 12 * gi);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const __half2 a = __floats2half2_rn(llr[4 * k], llr[4 * k + 1]), b = __floats2half2_rn(llr[4 * k + 2], llr[4 * k + 3]);
            o[k] = make_uint2(*reinterpret_cast<const unsigned *>(&a), *reinterpret_cast<const unsigned *>(&b));
        }
    } else {
        if (!live) return;
        float4 *o = reinterpret_cast<float4 *>(static_cast<float *>(out) + 12 * gi);
#pragma unroll
        for (int k = 0; k < 3; k++) o[k] = make_float4(llr[4 * k], llr[4 * k + 1], llr[4 * k + 2], llr[4 * k + 3]);
    }
}

template <int M, typename T>
cudaError_t demap_m(const DemapArgs &a, cudaStream_t st)
{
    const T *si = static_cast<const T *>(a.sym_i), *sq = static_cast<const T *>(a.sym_q);
    if (a.llr_type == TDB200_LLR_F64) {
        const size_t n_sym = a.n_llr / M;
        demap64_kernel<M, T><<<(unsigned)((n_sym + 255) / 256), 256, 0, st>>>(si, sq, static_cast<double *>(a.llr), n_sym, a.kf);
    } else {
        const size_t ng = a.n_llr / 12;
        const unsigned grid = (unsigned)((ng + 255) / 256);
        const float kf = (float)a.kf, scale = (float)(1 << a.frac_bits);
        // the 12-at-a-time kernel needs 16-byte aligned output groups; anything else goes symbol by symbol
        const size_t osz = a.llr_type == TDB200_LLR_S8 ? 1 : (a.llr_type == TDB200_LLR_F16 ? 2 : 4);
        const bool wide = ng > 0 && (reinterpret_cast<size_t>(a.llr) & (osz == 1 ? 3 : (osz == 2 ? 7 : 15))) == 0;
        if (wide) {
            if (a.llr_type == TDB200_LLR_S8) demap32_kernel<M, T, TDB200_LLR_S8><<<grid, 256, 0, st>>>(si, sq, a.llr, ng, kf, scale, a.clip);
            else if (a.llr_type == TDB200_LLR_F16) demap32_kernel<M, T, TDB200_LLR_F16><<<grid, 256, 0, st>>>(si, sq, a.llr, ng, kf, scale, a.clip);
            else demap32_kernel<M, T, TDB200_LLR_F32><<<grid, 256, 0, st>>>(si, sq, a.llr, ng, kf, scale, a.clip);
        }
        const size_t done = wide ? ng * 12 : 0;
        if (done < a.n_llr) {  // the tail (at most 11 values), or everything for an unaligned buffer
            const size_t n_sym = (a.n_llr - done) / M;
            demap32_tail_kernel<M, T><<<(unsigned)((n_sym + 255) / 256), 256, 0, st>>>(si, sq, a.llr, a.llr_type, done / M, n_sym, kf, scale, a.clip);
        }
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t demap_t(const DemapArgs &a, cudaStream_t st)
{
    switch (a.modulation) {
        case 1: return demap_m<1, T>(a, st);
        case 2: return demap_m<2, T>(a, st);
        case 3: return demap_m<3, T>(a, st);
        case 4: return demap_m<4, T>(a, st);
        case 6: return demap_m<6, T>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int M>
cudaError_t modulate_m(const uint8_t *coded, void *si, void *sq, int sym_type, size_t n_sym, cudaStream_t st)
{
    const unsigned grid = (unsigned)((n_sym + 255) / 256);
    if (sym_type == TDB200_LLR_F64) modulate_kernel<M, double><<<grid, 256, 0, st>>>(coded, static_cast<double *>(si), static_cast<double *>(sq), n_sym);
    else if (sym_type == TDB200_LLR_F16) modulate_kernel<M, __half><<<grid, 256, 0, st>>>(coded, static_cast<__half *>(si), static_cast<__half *>(sq), n_sym);
    else modulate_kernel<M, float><<<grid, 256, 0, st>>>(coded, static_cast<float *>(si), static_cast<float *>(sq), n_sym);
    return cudaGetLastError();
}

}  // namespace

bool modulation_ok(int M) { return M == 1 || M == 2 || M == 3 || M == 4 || M == 6; }

cudaError_t launch_modulate(const uint8_t *coded, void *si, void *sq, int sym_type, size_t n_bits, int M, cudaStream_t st)
{
    if (n_bits == 0) return cudaSuccess;
    const size_t n_sym = n_bits / M;
    switch (M) {
        case 1: return modulate_m<1>(coded, si, sq, sym_type, n_sym, st);
        case 2: return modulate_m<2>(coded, si, sq, sym_type, n_sym, st);
        case 3: return modulate_m<3>(coded, si, sq, sym_type, n_sym, st);
        case 4: return modulate_m<4>(coded, si, sq, sym_type, n_sym, st);
        case 6: return modulate_m<6>(coded, si, sq, sym_type, n_sym, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_awgn(const void *x, void *y, int type, size_t n, double sigma, unsigned long long seed, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)(((n + 3) / 4 + 255) / 256);
    if (type == TDB200_LLR_F64) awgn_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double *>(x), static_cast<double *>(y), n, (float)sigma, seed);
    else if (type == TDB200_LLR_F16) awgn_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half *>(x), static_cast<__half *>(y), n, (float)sigma, seed);
    else awgn_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(x), static_cast<float *>(y), n, (float)sigma, seed);
    return cudaGetLastError();
}

cudaError_t launch_demap(const DemapArgs &a, cudaStream_t st)
{
    if (a.n_llr == 0) return cudaSuccess;
    if (a.sym_type == TDB200_LLR_F64) return demap_t<double>(a, st);
    if (a.sym_type == TDB200_LLR_F16) return demap_t<__half>(a, st);
    return demap_t<float>(a, st);
}

}  // namespace tdb200
