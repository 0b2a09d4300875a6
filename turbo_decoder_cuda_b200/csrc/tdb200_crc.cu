// tdb200_crc.cu -- the 24-bit CRCs of the transport-block stage above the decode path (SURVEY.md 8f.3;
// TS 36.212 5.1.1: gCRC24A = 0x1864CFB for the transport block, gCRC24B = 0x1800063 for each code block
// of a segmented one).  The reference has only a placeholder here (previous/Decoder.cc:1026 "stoprule ...
// 1=CRC", :1098-1099).
//
// One warp per codeblock.  A CRC is linear over GF(2): crc(A || B) = crc(A) * x^|B| + crc(B) mod g, and
// leading zeros do not change it (zero initial state).  The n bits are right-aligned in 32 chunks of
// c = ceil(n/32) bits; every lane divides its own chunk -- eight bits per step where the chunk is 8-byte
// aligned (one 64-bit load of eight one-bit bytes, packed into a byte by two multiplies, then the byte-wise
// table of the generator, built in shared memory by the CTA), bit by bit at the ragged ends -- then five
// shuffle levels fold the 32 partial remainders together, multiplying the left half by x^(c*2^level)
// mod g (constants from the host) with a 24-step carry-less multiply.  `attach` writes the parity
// bits behind the first K-24 bits, `check` divides all K bits and reports remainder == 0.
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

__device__ __forceinline__ unsigned mulmod24(unsigned a, unsigned b, unsigned poly)
{
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) {
        r ^= (b & 1u) ? a : 0u;
        b >>= 1;
        a = ((a << 1) & 0xffffffu) ^ ((a & 0x800000u) ? poly : 0u);
    }
    return r;
}

__device__ __forceinline__ unsigned step1(unsigned crc, unsigned bit, unsigned poly)
{
    const unsigned fb = ((crc >> 23) & 1u) ^ (bit & 1u);
    return ((crc << 1) & 0xffffffu) ^ (fb ? poly : 0u);
}
// four one-bit bytes (first bit in the lowest byte) -> a nibble with the first bit on top: the four partial
// products of the multiply land on distinct bits, so nothing carries
__device__ __forceinline__ unsigned pack4(unsigned w) { return (((w & 0x01010101u) * 0x08040201u) >> 24) & 0xfu; }

__global__ void __launch_bounds__(128) crc24_kernel(CrcArgs A)
{
    __shared__ unsigned tab[256];  // tab[v] = v(x) * x^24 mod g
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        unsigned r = (unsigned)v << 16;
        for (int b = 0; b < 8; b++) r = ((r << 1) & 0xffffffu) ^ ((r & 0x800000u) ? A.poly : 0u);
        tab[v] = r;
    }
    __syncthreads();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= A.n_cb) return;
    uint8_t *row = A.bits + (size_t)warp * A.K;
    const int n = A.attach ? A.K - 24 : A.K;
    const int c = A.chunk, z = 32 * c - n;  // z leading virtual zeros
    unsigned crc = 0;
    int i = max(lane * c - z, 0);  // leading zeros leave the zero state alone
    const int end = lane * c - z + c;
    for (; i < end && (reinterpret_cast<size_t>(row + i) & 7); i++) crc = step1(crc, row[i], A.poly);
    for (; i + 8 <= end; i += 8) {
        const uint2 w = *reinterpret_cast<const uint2 *>(row + i);
        const unsigned byte = (pack4(w.x) << 4) | pack4(w.y);
        crc = ((crc << 8) & 0xffffffu) ^ tab[((crc >> 16) ^ byte) & 0xffu];
    }
    for (; i < end; i++) crc = step1(crc, row[i], A.poly);
#pragma unroll
    for (int lv = 0; lv < 5; lv++) {
        const int s = 1 << lv;
        const unsigned left = __shfl_up_sync(0xffffffffu, crc, s);
        const unsigned folded = mulmod24(left, A.xpow[lv], A.poly) ^ crc;
        if (((lane + 1) & (2 * s - 1)) == 0) crc = folded;
    }
    crc = __shfl_sync(0xffffffffu, crc, 31);
    if (A.attach) {
        if (lane < 24) row[A.K - 24 + lane] = (uint8_t)((crc >> (23 - lane)) & 1u);  // p_0 first (MSB)
    } else if (lane == 0) {
        if (A.ok) A.ok[warp] = (uint8_t)(crc == 0);
        if (A.remainder) A.remainder[warp] = (int32_t)crc;
    }
}

}  // namespace

// x^e mod g, g = x^24 + poly
static unsigned xpow_mod(long e, unsigned poly)
{
    unsigned a = 1;
    for (long i = 0; i < e; i++) a = ((a << 1) & 0xffffffu) ^ ((a & 0x800000u) ? poly : 0u);
    return a;
}

cudaError_t launch_crc24(const CrcArgs &a0, cudaStream_t st)
{
    if (a0.n_cb == 0) return cudaSuccess;
    CrcArgs a = a0;
    const int n = a.attach ? a.K - 24 : a.K;
    a.chunk = (n + 31) / 32;
    for (int lv = 0; lv < 5; lv++) a.xpow[lv] = xpow_mod((long)a.chunk << lv, a.poly);
    crc24_kernel<<<(a.n_cb + 3) / 4, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tdb200
