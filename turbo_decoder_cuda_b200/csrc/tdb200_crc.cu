// tdb200_crc.cu -- the 24-bit CRCs of the transport-block stage above the decode path (SURVEY.md 8f.3;
// TS 36.212 5.1.1: gCRC24A = 0x1864CFB for the transport block, gCRC24B = 0x1800063 for each code block
// of a segmented one).  The reference has only a placeholder here (previous/Decoder.cc:1026 "stoprule ...
// 1=CRC", :1098-1099).
//
// One warp per codeblock.  A CRC is linear over GF(2): crc(A || B) = crc(A) * x^|B| + crc(B) mod g, and
// leading zeros do not change it (zero initial state).  The n bits are right-aligned in 32 chunks of
// c = ceil(n/32) bits; every lane runs the bit-serial division over its own chunk, then five
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

__global__ void __launch_bounds__(128) crc24_kernel(CrcArgs A)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= A.n_cb) return;
    uint8_t *row = A.bits + (size_t)warp * A.K;
    const int n = A.attach ? A.K - 24 : A.K;
    const int c = A.chunk, z = 32 * c - n;  // z leading virtual zeros
    unsigned crc = 0;
    for (int i = 0; i < c; i++) {
        const int idx = lane * c + i - z;
        const unsigned bit = idx >= 0 ? (row[idx] & 1u) : 0u;
        const unsigned fb = ((crc >> 23) & 1u) ^ bit;
        crc = ((crc << 1) & 0xffffffu) ^ (fb ? A.poly : 0u);
    }
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
