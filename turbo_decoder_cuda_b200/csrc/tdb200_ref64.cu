// tdb200_ref64.cu -- fp64 reference-order Log-MAP decoder (TDB200_ALGO_LOGMAP_F64).
//
// The LLR-parity mode: an unsegmented BCJR that performs the reference's floating-point
// operations on the same operands in the same order as
//     TurboDecoding()    ITTC/log_map.cpp:1146-1280
//     Log_MAP_decoder()  ITTC/log_map.cpp:898-1047   (gamma :962-972, alpha :975-1001,
//                                                      beta :1004-1021, LLR :1024-1039)
//     E_algorithm()      ITTC/log_map.cpp:779-801, LUT :14-18
// so its a-posteriori/extrinsic LLRs agree with the CPU code to rounding noise (the only
// non-reproducible term in the reference is its uninitialised tempmax[], :925/:989; like the
// oracle this kernel normalises by max_j alpha_j).
//
// Mapping (nothing like the reference's loops): one warp owns four codeblocks for the whole
// decode.  The two sequential recursions run with ONE LANE PER TRELLIS STATE (8 lanes per
// codeblock, 4 codeblocks per warp); predecessor/successor metrics are exchanged with
// width-8 warp shuffles, the per-step max_j alpha_j is a 3-round shuffle butterfly, and the
// (xs,xp,La) inputs of eight consecutive steps are fetched with one coalesced load per lane and
// broadcast by shuffle.  alpha/beta go to an HBM workspace ([step][state], 64 B per step); the
// LLR/extrinsic phase is then embarrassingly parallel over trellis positions and runs with all
// 32 lanes.  Warps never talk to each other, so the only synchronisation is __syncwarp().
// Compile with -fmad=false: products here are exact (x * +-1, x * 0.5) so contraction would not
// change results, but the flag keeps that a non-question.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

constexpr double kInfty = 1E20;  // ITTC/log_map.h:72-74

// max*(x,y), ITTC/log_map.cpp:779-801.  The linear LUT scan is restated as a 4-level select
// tree over the same 16 breakpoints: region [idx[k], idx[k+1]) -> table[k], d >= 4.3758 -> 0.
__device__ __forceinline__ double lut_corr(double d)
{
    return d < 1.0502
               ? (d < 0.43275 ? (d < 0.19587 ? (d < 0.08824 ? 0.69315 : 0.65) : (d < 0.31026 ? 0.6 : 0.55))
                              : (d < 0.70963 ? (d < 0.56508 ? 0.5 : 0.45) : (d < 0.86972 ? 0.4 : 0.35)))
               : (d < 2.2522 ? (d < 1.5078 ? (d < 1.2587 ? 0.3 : 0.25) : (d < 1.8212 ? 0.2 : 0.15))
                             : (d < 3.6764 ? (d < 2.9706 ? 0.1 : 0.05) : (d < 4.3758 ? 0.025 : 0.0)));
}
__device__ __forceinline__ double max_star(double x, double y)
{
    double d = (y - x) > 0 ? (y - x) : (x - y);
    return (x > y ? x : y) + lut_corr(d);
}

__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

__device__ __forceinline__ double load_llr_half(const void *p, int type, size_t idx)
{
    // flow_for_decode[i] *= 0.5, ITTC/log_map.cpp:1202-1205 (done on a copy)
    if (type == TDB200_LLR_F64) return static_cast<const double *>(p)[idx] * 0.5;
    if (type == TDB200_LLR_F32) return static_cast<double>(static_cast<const float *>(p)[idx]) * 0.5;
    if (type == TDB200_LLR_F16) return static_cast<double>(__half2float(static_cast<const __half *>(p)[idx])) * 0.5;
    return static_cast<double>(static_cast<const int8_t *>(p)[idx]) * 0.0625;  // S8, 3 fractional bits
}

// One BCJR pass for the four codeblocks of this warp.  xs/xp/La/LLR/tmax/alpha/beta point at
// the warp's first codeblock; strides are per codeblock.
struct SisoPtrs {
    const double *xs, *xp, *La;
    double *LLR, *tmax, *alpha, *beta;
    size_t sT;   // stride of T-long arrays
    size_t sT1;  // stride of tmax (T+1)
    size_t sAB;  // stride of alpha/beta (8*(T+1))
};

__device__ void siso_pass(const SisoPtrs &p, int T, int terminated, int lane)
{
    const int g = lane >> 3, j = lane & 7;
    const double *xs = p.xs + g * p.sT, *xp = p.xp + g * p.sT, *La = p.La + g * p.sT;
    double *tmax = p.tmax + g * p.sT1;
    double *alpha = p.alpha + g * p.sAB, *beta = p.beta + g * p.sAB;
    const int ls0 = tb(kLs0, j), ls1 = tb(kLs1, j), ns0 = tb(kNs0, j), ns1 = tb(kNs1, j);
    const double sa0 = o0(ls0), sa1 = o1(ls1);  // parity signs of the branches ENTERING state j
    const double sb0 = o0(j), sb1 = o1(j);      // parity signs of the branches LEAVING state j

    // ---- alpha forward, :975-1001
    double al = (j == 0) ? 0.0 : -kInfty;  // :943-948
    alpha[j] = al;
    for (int i0 = 0; i0 < T; i0 += 8) {
        const int ix = i0 + j;
        const double vs = ix < T ? xs[ix] : 0.0, vp = ix < T ? xp[ix] : 0.0, vl = ix < T ? La[ix] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = i0 + u;
            if (i >= T) break;  // uniform
            const double s = shfl8(vs, u), q = shfl8(vp, u), l = shfl8(vl, u);
            const double g0 = -s + q * sa0 - l / 2;  // gamma(from ls0, input 0), :967-968
            const double g1 = s + q * sa1 + l / 2;   // gamma(from ls1, input 1), :969-970
            const double tx = g0 + shfl8(al, ls0);
            const double ty = g1 + shfl8(al, ls1);
            const double v = max_star(tx, ty);
            double m = v;  // tempmax[i+1] = max_j alpha_j, :987-993
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const double t = __shfl_xor_sync(0xffffffffu, m, o, 8);
                m = (m < t) ? t : m;
            }
            // the reference compares against an uninitialised tempmax[] (:925,:989); on a clean
            // (zero-filled) heap that is max(0, max_j alpha_j), which is what oracle/ pins
            m = (m < 0.0) ? 0.0 : m;
            al = v - m;  // :996-999
            alpha[(size_t)(i + 1) * 8 + j] = al;
            if (j == 0) tmax[i + 1] = m;
        }
    }
    __syncwarp();
    // ---- beta backward, :1004-1021
    double be = (j == 0) ? 0.0 : (terminated ? -kInfty : 0.0);  // :944-959
    beta[(size_t)T * 8 + j] = be;
    for (int i0 = ((T - 1) >> 3) << 3; i0 >= 0; i0 -= 8) {
        const int ix = i0 + j;
        const double vs = ix < T ? xs[ix] : 0.0, vp = ix < T ? xp[ix] : 0.0, vl = ix < T ? La[ix] : 0.0;
        const double vm = ix < T ? tmax[ix + 1] : 0.0;
#pragma unroll
        for (int u = 7; u >= 0; u--) {
            const int i = i0 + u;
            if (i >= T) continue;  // uniform
            const double s = shfl8(vs, u), q = shfl8(vp, u), l = shfl8(vl, u), m = shfl8(vm, u);
            const double g0 = -s + q * sb0 - l / 2;
            const double g1 = s + q * sb1 + l / 2;
            const double tx = g0 + shfl8(be, ns0);
            const double ty = g1 + shfl8(be, ns1);
            be = max_star(tx, ty) - m;
            beta[(size_t)i * 8 + j] = be;
        }
    }
    __syncwarp();
    // ---- LLR, :1024-1039: all 32 lanes over positions, one codeblock after the other
    for (int c = 0; c < 4; c++) {
        const double *cxs = p.xs + c * p.sT, *cxp = p.xp + c * p.sT, *cLa = p.La + c * p.sT;
        const double *ca = p.alpha + c * p.sAB, *cb = p.beta + c * p.sAB;
        double *cL = p.LLR + c * p.sT;
        for (int i = lane; i < T; i += 32) {
            const double s = cxs[i], q = cxp[i], l = cLa[i];
            double a[8], b[8];
            const double2 *a2 = reinterpret_cast<const double2 *>(ca + (size_t)i * 8);
            const double2 *b2 = reinterpret_cast<const double2 *>(cb + (size_t)(i + 1) * 8);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double2 t = a2[k];
                a[2 * k] = t.x; a[2 * k + 1] = t.y;
                t = b2[k];
                b[2 * k] = t.x; b[2 * k + 1] = t.y;
            }
            double m0 = 0, m1 = 0;
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int l0 = tb(kLs0, jj), l1 = tb(kLs1, jj);
                const double t0 = (-s + q * o0(l0) - l / 2) + a[l0] + b[jj];  // :1028-1030
                const double t1 = (s + q * o1(l1) + l / 2) + a[l1] + b[jj];   // :1032-1034
                if (jj == 0) { m0 = t0; m1 = t1; }
                else { m0 = max_star(m0, t0); m1 = max_star(m1, t1); }        // E_algorithm_seq, :817-829
            }
            cL[i] = m1 - m0;  // :1038
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) ref64_decode_kernel(Ref64Args a)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int cb0 = warp * 4;
    if (cb0 >= a.n_cb) return;
    const int K = a.K, T = K + kTail, NL = 3 * K + 4 * kTail;
    const size_t sT = T, sT1 = T + 1, sAB = (size_t)8 * (T + 1);
    const Ref64Workspace &w = a.ws;
    double *xs1 = w.xs1 + cb0 * sT, *xp1 = w.xp1 + cb0 * sT, *xs2 = w.xs2 + cb0 * sT, *xp2 = w.xp2 + cb0 * sT;
    double *La = w.La + cb0 * sT, *Le = w.Le + cb0 * sT, *LLR = w.LLR + cb0 * sT;

    // ---- x0.5 and demultiplex, :1202-1209, :1083-1127
    for (int c = 0; c < 4; c++) {
        const int cb = cb0 + c;
        const bool valid = cb < a.n_cb;
        const size_t base = (size_t)(valid ? cb : cb0) * NL;
        for (int i = lane; i < T; i += 32) {
            double s1, p1, s2, p2;
            if (i < K) {
                s1 = load_llr_half(a.llr, a.llr_type, base + 3 * i);
                p1 = load_llr_half(a.llr, a.llr_type, base + 3 * i + 1);
                p2 = load_llr_half(a.llr, a.llr_type, base + 3 * i + 2);
                s2 = load_llr_half(a.llr, a.llr_type, base + 3 * (size_t)a.pi[i]);
            } else {
                const int m = i - K;
                s1 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * m);
                p1 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * m + 1);
                s2 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * kTail + 2 * m);
                p2 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * kTail + 2 * m + 1);
            }
            xs1[c * sT + i] = s1; xp1[c * sT + i] = p1; xs2[c * sT + i] = s2; xp2[c * sT + i] = p2;
            Le[c * sT + i] = 0.0;  // :1212-1215
        }
    }
    __syncwarp();

    SisoPtrs sp;
    sp.La = La; sp.LLR = LLR;
    sp.tmax = w.tmax + cb0 * sT1; sp.alpha = w.alpha + cb0 * sAB; sp.beta = w.beta + cb0 * sAB;
    sp.sT = sT; sp.sT1 = sT1; sp.sAB = sAB;

    for (int it = 0; it < a.n_iter; it++) {
        const bool last = (it == a.n_iter - 1);
        for (int siso = 0; siso < 2; siso++) {
            // a-priori for this pass: SISO-1 La[pi(i)] = Le[i] (random_deinterlvr_double, :1221),
            // SISO-2 La[i] = Le[pi(i)] (randominterleaver_double, :1242); tail forced to 0.
            const int *idx = siso == 0 ? a.pi_inv : a.pi;
            for (int c = 0; c < 4; c++)
                for (int i = lane; i < T; i += 32) La[c * sT + i] = (i < K) ? Le[c * sT + idx[i]] : 0.0;
            __syncwarp();
            sp.xs = siso == 0 ? xs1 : xs2;
            sp.xp = siso == 0 ? xp1 : xp2;
            siso_pass(sp, T, 1, lane);
            // extrinsic, :1234-1238 / :1255-1259; decision + deinterleave, :1261-1264
            for (int c = 0; c < 4; c++) {
                const int cb = cb0 + c;
                const bool valid = cb < a.n_cb;
                for (int i = lane; i < T; i += 32) {
                    const double L = LLR[c * sT + i];
                    Le[c * sT + i] = L - La[c * sT + i] - 2 * sp.xs[c * sT + i];
                    if (!valid) continue;
                    if (siso == 1 && i < K) {
                        const int bit = (L < 0) ? 0 : 1;
                        const int pos = a.pi[i];
                        if (a.bits_iters) a.bits_iters[((size_t)cb * a.n_iter + it) * K + pos] = bit;
                        if (last && a.bits) a.bits[(size_t)cb * K + pos] = (uint8_t)bit;
                    }
                    if (last) {
                        if (siso == 0 && a.llr1) a.llr1[(size_t)cb * T + i] = L;
                        if (siso == 1 && a.llr2) a.llr2[(size_t)cb * T + i] = L;
                        if (siso == 1 && a.ext2) a.ext2[(size_t)cb * T + i] = Le[c * sT + i];
                    }
                }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(128) ref64_siso_kernel(Ref64SisoArgs a)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int cb0 = warp * 4;
    if (cb0 >= a.n_cb) return;
    const int T = a.T;
    const size_t sT = T, sT1 = T + 1, sAB = (size_t)8 * (T + 1);
    const Ref64Workspace &w = a.ws;
    // de-interleave recs (xs,xp pairs) into the planar workspace
    for (int c = 0; c < 4; c++) {
        const int cb = min(cb0 + c, a.n_cb - 1);
        for (int i = lane; i < T; i += 32) {
            w.xs1[(cb0 + c) * sT + i] = a.recs[(size_t)cb * 2 * T + 2 * i];
            w.xp1[(cb0 + c) * sT + i] = a.recs[(size_t)cb * 2 * T + 2 * i + 1];
            w.La[(cb0 + c) * sT + i] = a.La[(size_t)cb * T + i];
        }
    }
    __syncwarp();
    SisoPtrs sp;
    sp.xs = w.xs1 + cb0 * sT; sp.xp = w.xp1 + cb0 * sT; sp.La = w.La + cb0 * sT; sp.LLR = w.LLR + cb0 * sT;
    sp.tmax = w.tmax + cb0 * sT1; sp.alpha = w.alpha + cb0 * sAB; sp.beta = w.beta + cb0 * sAB;
    sp.sT = sT; sp.sT1 = sT1; sp.sAB = sAB;
    siso_pass(sp, T, a.terminated, lane);
    for (int c = 0; c < 4; c++) {
        const int cb = cb0 + c;
        if (cb >= a.n_cb) break;
        for (int i = lane; i < T; i += 32) a.LLR[(size_t)cb * T + i] = w.LLR[(cb0 + c) * sT + i];
    }
}

}  // namespace

cudaError_t launch_ref64_decode(const Ref64Args &a, cudaStream_t st, int *n_launches)
{
    const int warps = (a.n_cb + 3) / 4;
    const int wpb = 2;  // 64-thread CTAs: spreads a small batch over more SMs
    const int blocks = (warps + wpb - 1) / wpb;
    ref64_decode_kernel<<<blocks, wpb * 32, 0, st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ref64_siso(const Ref64SisoArgs &a, cudaStream_t st, int *n_launches)
{
    const int warps = (a.n_cb + 3) / 4;
    const int wpb = 2;
    const int blocks = (warps + wpb - 1) / wpb;
    ref64_siso_kernel<<<blocks, wpb * 32, 0, st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

}  // namespace tdb200
