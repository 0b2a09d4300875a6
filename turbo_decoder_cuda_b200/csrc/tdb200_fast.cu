// tdb200_fast.cu -- host side of TDB200_ALGO_MAXLOG_S16, the throughput decoder: kernel selection by
// geometry and channel-LLR type, shared-memory sizing, launch.  The device code (and the description
// of the algorithm) is in tdb200_fast_kernel.cuh, instantiated per LLR type in tdb200_fast_inst_*.cu.
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {

typedef void (*fast_kernel_fn)(FastArgs);
// one translation unit per channel-LLR type (tdb200_fast_inst_*.cu, all from tdb200_fast_kernel.cuh)
fast_kernel_fn fast_pick_f64(const FastGeom &g);
fast_kernel_fn fast_pick_f32(const FastGeom &g);
fast_kernel_fn fast_pick_s8(const FastGeom &g);
fast_kernel_fn fast_pick_f16(const FastGeom &g);
// the same with the CRC stopping rule compiled in (tdb200_fast_inst_crc_*.cu)
fast_kernel_fn fast_pick_crc_f64(const FastGeom &g);
fast_kernel_fn fast_pick_crc_f32(const FastGeom &g);
fast_kernel_fn fast_pick_crc_s8(const FastGeom &g);
fast_kernel_fn fast_pick_crc_f16(const FastGeom &g);

// the Log-MAP variants (tdb200_fast_inst_lm_*.cu)
fast_kernel_fn fast_pick_lm_f64(const FastGeom &g);
fast_kernel_fn fast_pick_lm_f32(const FastGeom &g);
fast_kernel_fn fast_pick_lm_s8(const FastGeom &g);
fast_kernel_fn fast_pick_lm_f16(const FastGeom &g);

// received BPSK / QPSK float symbols, demapped in the load stage (tdb200_fast_inst_sym*.cu, tdb200_fast_inst_lm_sym*.cu)
fast_kernel_fn fast_pick_sym1(const FastGeom &g);
fast_kernel_fn fast_pick_sym2(const FastGeom &g);
fast_kernel_fn fast_pick_lm_sym1(const FastGeom &g);
fast_kernel_fn fast_pick_lm_sym2(const FastGeom &g);

namespace {

fast_kernel_fn pick_kernel(const FastGeom &g, int llr_type, bool crc = false, bool logmap = false)
{
    if (llr_type == kLlrSymBpskF32) return logmap ? fast_pick_lm_sym1(g) : fast_pick_sym1(g);
    if (llr_type == kLlrSymQpskF32) return logmap ? fast_pick_lm_sym2(g) : fast_pick_sym2(g);
    if (logmap) {
        switch (llr_type) {
            case TDB200_LLR_F32: return fast_pick_lm_f32(g);
            case TDB200_LLR_F64: return fast_pick_lm_f64(g);
            case TDB200_LLR_F16: return fast_pick_lm_f16(g);
            default: return fast_pick_lm_s8(g);
        }
    }
    if (crc) {
        switch (llr_type) {
            case TDB200_LLR_F32: return fast_pick_crc_f32(g);
            case TDB200_LLR_F64: return fast_pick_crc_f64(g);
            case TDB200_LLR_F16: return fast_pick_crc_f16(g);
            default: return fast_pick_crc_s8(g);
        }
    }
    switch (llr_type) {
        case TDB200_LLR_F32: return fast_pick_f32(g);
        case TDB200_LLR_F64: return fast_pick_f64(g);
        case TDB200_LLR_F16: return fast_pick_f16(g);
        default: return fast_pick_s8(g);
    }
}

}  // namespace

bool fast_s16_specialised(const FastGeom &g, bool logmap)
{
    if (logmap) return fast_spec_lm(g);
    return fast_spec_pn(g) || fast_spec128g8(g) || fast_spec192(g) || fast_spec_rt(g) || fast_spec_rt192(g);
}

// bytes of one codeblock-pair region / of the part shared by the pairs of a CTA of `threads` threads
int fast_s16_pair_bytes(const FastGeom &g)
{
    const int W = g.L * g.PP, Wp = (W + 7) & ~7;
    int b = 3 * 4 * Wp + 2 * Wp + 4 * g.n_ckpt * 7 * g.P + 4 * ((g.NW + 1) / 2) * g.P + 4 * 16;  // ..., decisions, tail vectors
    b = (b + 3) & ~3;
    // Pairs that share a CTA sit side by side; thread (pair q, sub-block t) reads word q * stride + j * PP + t.
    // With stride = P (mod 32 words) that is bank tid + const: conflict-free across the pairs of a warp (a stride
    // that is a multiple of 16 words put K = 64 on 4 banks).
    if (g.P <= 64)
        while (((b / 4) & 31) != (g.P & 31)) b += 4;
    return b;
}
static int shared_bytes(const FastGeom &g, int threads, int np)
{
    const int W = g.L * g.PP, Wp = (W + 7) & ~7;
    return 2 * Wp + 4 * 32 * (threads / 32) + 4 * ((np + 3) & ~3);  // table, warp-edge words, stop flags
}
int fast_s16_smem_bytes(const FastGeom &g) { return ((g.pair_bytes * g.NP + 15) & ~15) + shared_bytes(g, g.threads, g.NP); }

cudaError_t fast_s16_configure(FastGeom &g, int sm_count, bool logmap)
{
    // The attribute belongs to the kernel, not to a decoder handle: several handles with
    // different geometries share one instantiation, so always opt in to the device maximum.
    int dev = 0, optin = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e0 != cudaSuccess) return e0;
    for (int t : {(int)TDB200_LLR_F64, (int)TDB200_LLR_F32, (int)TDB200_LLR_S8, (int)TDB200_LLR_F16, kLlrSymBpskF32, kLlrSymQpskF32}) {
        cudaError_t e = cudaFuncSetAttribute(pick_kernel(g, t, false, logmap), cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e == cudaSuccess && !logmap && t <= TDB200_LLR_F16) e = cudaFuncSetAttribute(pick_kernel(g, t, true), cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e0 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pick_kernel(g, TDB200_LLR_F32, false, logmap), g.threads, g.smem_bytes);
    if (e0 != cudaSuccess) return e0;
    g.resident_ctas = per_sm * sm_count;
    return cudaSuccess;
}

cudaError_t launch_fast_s16(const FastArgs &a0, cudaStream_t st, int *n_launches)
{
    FastArgs a = a0;
    const int pairs = (a.n_cb + 1) / 2;
    // several pairs per CTA when a codeblock needs few threads -- but not so many that a small batch
    // leaves SMs without work
    int np = a.g.NP;
    while (np > 1 && (pairs + np - 1) / np < 2 * a.sm_count) np = (np + 1) / 2;
    a.pairs_per_cta = np;
    const int threads = ((np * a.g.P + 31) / 32) * 32;
    const int smem = ((a.g.pair_bytes * np + 15) & ~15) + shared_bytes(a.g, threads, np);
    pick_kernel(a.g, a.llr_type, a.early_term == 2, a.logmap != 0)<<<(pairs + np - 1) / np, threads, smem, st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

}  // namespace tdb200
