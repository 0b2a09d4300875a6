// tdb200_encode.cu -- the caller side of the decode path, on the device:
//
//   encode_kernel   TurboEnCoding() -> encoderm_turbo() -> rsc_encode()   ITTC/log_map.cpp:700-730, 530-583, 451-527
//                   (13,15)_8 PCCC with trellis termination, the reference's multiplex order
//                   [3i]=x_i, [3i+1]=z_i, [3i+2]=z'_i, then (x,z)x3 of RSC1 and (x',z')x3 of RSC2 (:566-578)
//   channel_kernel  module() BPSK + AWGN() + demodule()                   ITTC/main.cpp:197-202, modanddem.cpp:175-224
//                   LLR = 2 r / sigma^2 with r = (2c-1) + sigma*n.  The reference draws n from a 12-term
//                   central-limit sum over rand() (mgrns, log_map.cpp:1359-1392, seeded from time());
//                   here n is Philox4x32-10 + Box-Muller, a pure function of (seed, element index).
//
// The encoder is a linear recursion over GF(2), a_k = d_k ^ a_{k-2} ^ a_{k-3}; nothing like the
// reference's bit-serial loop is needed to parallelise it.  One warp encodes one codeblock: lane l owns
// a chunk of ceil(K/32) consecutive trellis steps.  Pass 1 runs the chunk from state 0 together
// with the three unit states under zero input (four recursions packed in the bits of one word),
// which yields the chunk's zero-state response z_l and its 3x3 state-transition matrix T_l.  The
// true entry state of every chunk follows from the 32-step scan s_{l+1} = T_l s_l ^ z_l across the
// lanes (shuffles); pass 2 re-runs the chunk from that state and emits the parity bits.  RSC2 reads
// its input through the QPP table.  Bits, table and the coded row live in shared memory in a padded
// chunk layout (odd word stride between the lanes' chunks: no bank conflicts); global traffic is
// 8-byte loads and 4-byte stores, coalesced.  Measured on B200: 0.18 ms for 4096 codeblocks of K = 6144
// (561 GB/s of input + output bytes; the serial recursion inside a lane, not memory, is the limit).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

// One trellis step on four packed recursions (bit j of every word belongs to recursion j).
// Registers s0 (newest), s1, s2; feedback 1011, forward 1101 (ITTC/log_map.h:34-36).
__device__ __forceinline__ unsigned rsc_step4(unsigned d, unsigned &s0, unsigned &s1, unsigned &s2)
{
    const unsigned a = d ^ s1 ^ s2;
    const unsigned p = a ^ s0 ^ s2;
    s2 = s1; s1 = s0; s0 = a;
    return p;
}

// Shared-memory position of byte i of an array that is cut into chunks of `chunk` bytes, one chunk per
// lane: chunks are `stride` bytes apart with stride/4 odd, so the 32 lanes, each walking its own
// chunk, always hit 32 different banks.
// `magic` = ceil(2^32 / chunk): i / chunk == mulhi(i, magic) for the index ranges here (i, chunk < 2^16).
__device__ __forceinline__ int padpos(int i, int chunk, unsigned magic, int stride)
{
    const int c = (int)__umulhi((unsigned)i, magic);
    return c * stride + (i - c * chunk);
}

__global__ void __launch_bounds__(128) encode_kernel(EncodeArgs A)
{
    // shared memory: the QPP table as padded bit positions (uint16, one copy per CTA), then per warp the
    // K input bits and the coded row under construction, both in the padded chunk layout -- every
    // global access of this kernel is a coalesced vector access, every shared access conflict-free
    extern __shared__ __align__(16) unsigned char stage_raw[];
    const int K = A.K;
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + wic;
    const bool valid = warp < A.n_cb;
    const int NL = 3 * K + 12;      // a multiple of 4: rows leave with 32-bit stores
    const int C = (K + 31) / 32;    // trellis steps per lane
    const int SB = A.stride_bits, SO = A.stride_out, SP = A.stride_pi;  // chunk strides: bits / coded row (bytes), table (entries)
    const int bits_bytes = 32 * SB, out_bytes = 33 * SO;                // 33rd chunk: the 12 tail bytes
    uint16_t *spi = reinterpret_cast<uint16_t *>(stage_raw);
    uint8_t *sbits = stage_raw + ((2 * 32 * SP + 15) & ~15) + (size_t)wic * (bits_bytes + out_bytes);
    uint8_t *out = sbits + bits_bytes;
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        spi[padpos(i, C, A.magic_c, SP)] = (uint16_t)padpos(__ldg(A.pi + i), C, A.magic_c, SB);
    if (valid) {
        const uint2 *src = reinterpret_cast<const uint2 *>(A.bits + (size_t)warp * K);  // K is a multiple of 8
        for (int q = lane; q < K / 8; q += 32) {
            const uint2 v = __ldg(src + q);
#pragma unroll
            for (int j = 0; j < 8; j++) sbits[padpos(8 * q + j, C, A.magic_c, SB)] = (uint8_t)(((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 0xffu);
        }
    }
    __syncthreads();
    if (!valid) return;
    const int lo = min(lane * C, K), hi = min(lo + C, K);
    const uint8_t *my_bits = sbits + lane * SB - lo;   // my_bits[i] = bit i for i in [lo, hi)
    uint8_t *my_out = out + lane * SO - 3 * lo;        // my_out[3i+j] = coded byte 3i+j
    const uint16_t *my_pi = spi + lane * SP - lo;      // my_pi[i] = padded position of bit pi(i)

    for (int enc = 0; enc < 2; enc++) {
        // ---- pass 1: zero-state response of the chunk (bit 0) and images of the unit states (bits 1..3)
        unsigned s0 = 2u, s1 = 4u, s2 = 8u;  // recursion j+1 starts in unit state e_j (s0, s1, s2)
#pragma unroll 4
        for (int i = lo; i < hi; i++) {
            const unsigned d = (enc ? sbits[my_pi[i]] : my_bits[i]) & 1u;
            rsc_step4(d, s0, s1, s2);
        }
        // ---- scan over the lanes: entry state of lane l (3 bits: s0 | s1<<1 | s2<<2)
        unsigned st = 0;  // lane 0 starts in the all-zero state
        for (int l = 0; l < 31; l++) {
            // state after lane l's chunk = T_l * st ^ z_l, evaluated by lane l, handed to lane l+1
            unsigned nx0 = (s0 & 1u), nx1 = (s1 & 1u), nx2 = (s2 & 1u);
#pragma unroll
            for (int j = 0; j < 3; j++)
                if ((st >> j) & 1u) { nx0 ^= (s0 >> (j + 1)) & 1u; nx1 ^= (s1 >> (j + 1)) & 1u; nx2 ^= (s2 >> (j + 1)) & 1u; }
            const unsigned nxt = nx0 | (nx1 << 1) | (nx2 << 2);
            const unsigned got = __shfl_sync(0xffffffffu, nxt, l);
            if (lane == l + 1) st = got;
        }
        // ---- pass 2: the chunk from its true entry state, parity out
        unsigned r0 = st & 1u, r1 = (st >> 1) & 1u, r2 = (st >> 2) & 1u;
#pragma unroll 4
        for (int i = lo; i < hi; i++) {
            const unsigned d = (enc ? sbits[my_pi[i]] : my_bits[i]) & 1u;
            const unsigned p = rsc_step4(d, r0, r1, r2) & 1u;
            if (enc == 0) { my_out[3 * i] = (uint8_t)d; my_out[3 * i + 1] = (uint8_t)p; }
            else my_out[3 * i + 2] = (uint8_t)p;
        }
        // ---- termination (rsc_encode :483-491): the lane holding the end of the block runs three more steps
        //      with d = s1 ^ s2, which drives the feedback sum to zero
        const int last = (K - 1) / C;
        if (lane == last) {
            uint8_t *tail = out + 32 * SO;  // the 33rd chunk
            for (int m = 0; m < 3; m++) {
                const unsigned d = (r1 ^ r2) & 1u;
                const unsigned p = rsc_step4(d, r0, r1, r2) & 1u;
                tail[6 * enc + 2 * m] = (uint8_t)d;
                tail[6 * enc + 2 * m + 1] = (uint8_t)p;
            }
        }
    }
    // ---- the finished row leaves with coalesced 32-bit stores (bytes gathered out of the padded layout)
    __syncwarp();
    uint32_t *g = reinterpret_cast<uint32_t *>(A.coded + (size_t)warp * NL);
    for (int q = lane; q < NL / 4; q += 32) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = 4 * q + j;
            const int pos = o < 3 * K ? padpos(o, 3 * C, A.magic_3c, SO) : 32 * SO + (o - 3 * K);
            w |= (uint32_t)out[pos] << (8 * j);
        }
        g[q] = w;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) channel_kernel(ChannelArgs A, T *llr)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // four elements per thread
    const size_t i0 = 4 * q;
    if (i0 >= A.n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(A.seed, /*subsequence*/ q, /*offset*/ 0, &st);
    const float4 g = curand_normal4(&st);
    const float nz[4] = {g.x, g.y, g.z, g.w};
    const float k = 2.0f / (A.sigma * A.sigma);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const size_t i = i0 + j;
        if (i < A.n) {
            const float x = A.coded[i] ? 1.0f : -1.0f;  // module(): bit 1 -> +1 (positive LLR = bit 1)
            llr[i] = static_cast<T>((x + A.sigma * nz[j]) * k);
        }
    }
}

}  // namespace

// chunk stride in bytes for chunks of n bytes: the smallest multiple of 4 >= n whose word count is odd
static int odd_word_stride(int n) { return 4 * (((n + 3) / 4) | 1); }

cudaError_t launch_encode(const EncodeArgs &a0, cudaStream_t st)
{
    if (a0.n_cb == 0) return cudaSuccess;
    EncodeArgs a = a0;
    const int warps_per_cta = 4;
    const int C = (a.K + 31) / 32;
    a.stride_bits = odd_word_stride(C);            // bytes
    a.stride_out = odd_word_stride(3 * C);         // bytes
    a.stride_pi = odd_word_stride(2 * C) / 2;      // uint16 entries
    a.magic_c = (unsigned)((0x100000000ull + C - 1) / C);
    a.magic_3c = (unsigned)((0x100000000ull + 3 * C - 1) / (3 * C));
    const int smem = ((2 * 32 * a.stride_pi + 15) & ~15) + warps_per_cta * (32 * a.stride_bits + 33 * a.stride_out);  // 114 KB at K = 6144
    static bool configured[64] = {};  // the attribute is per device; setting it twice from two threads is harmless
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    encode_kernel<<<(a.n_cb + warps_per_cta - 1) / warps_per_cta, 32 * warps_per_cta, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_channel(const ChannelArgs &a, void *llr, int llr_type, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    const size_t threads = (a.n + 3) / 4;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (llr_type == TDB200_LLR_F64) channel_kernel<double><<<grid, 256, 0, st>>>(a, static_cast<double *>(llr));
    else if (llr_type == TDB200_LLR_F16) channel_kernel<__half><<<grid, 256, 0, st>>>(a, static_cast<__half *>(llr));
    else channel_kernel<float><<<grid, 256, 0, st>>>(a, static_cast<float *>(llr));
    return cudaGetLastError();
}

}  // namespace tdb200
