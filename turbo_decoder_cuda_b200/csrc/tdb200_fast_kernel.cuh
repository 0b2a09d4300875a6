// tdb200_fast_kernel.cuh -- TDB200_ALGO_MAXLOG_S16: the throughput decoder.
//
// What it computes (bit-exact integer specification: oracle/turbo_oracle_fx.c):
// the iterative PCCC decode of TurboDecoding() (ITTC/log_map.cpp:1146-1280) with the
// component decoder Log_MAP_decoder() (:898-1047) evaluated as max-log-MAP in 16-bit fixed
// point.  How it is laid out has nothing in common with the reference's loops:
//
//  * ONE CTA DECODES TWO CODEBLOCKS, ALL ITERATIONS, ON CHIP.  The two codeblocks ride in the
//    low/high halves of every 32-bit register (s16x2), so each VIADD.16x2 / VIADDMNMX.S16x2
//    advances both.  Channel LLRs are read from HBM exactly once (128-bit loads, fused with
//    quantisation and de-multiplexing into shared memory, cf. demultiplex() :1083-1127) and only
//    hard decisions go back.
//  * SUB-BLOCK PARALLEL BCJR.  The K-step trellis is cut into P = K/L sub-blocks; thread t
//    owns steps [tL,(t+1)L) and keeps all 8 state metrics in registers (no shuffles, no
//    barriers inside the recursions).  Boundary metrics come from the neighbouring sub-block:
//    the vector it saved G steps before the boundary in the PREVIOUS iteration (next-iteration
//    initialisation) is re-run over those G guard steps (warm-up) before each pass.
//  * ALPHA IS RECOMPUTED, NOT STORED.  A forward sweep leaves one alpha checkpoint per 8-step
//    window (7 words, in shared memory); the backward sweep re-creates the 8 alpha vectors of a
//    window in registers, then runs beta and the extrinsic output over it.  Shared memory
//    therefore holds only the a-priori/parity values, never the 8 x K metric array.
//  * BRANCH METRICS ARE FREE.  With gamma(b,c) = b*U + c*V (U = Ls + La, V = Lp; the per-step
//    constant the reference adds to every branch is dropped) a trellis step is 5 adds + 8 fused
//    add-max; the reference's gamma table (:962-972) never exists.
//  * U IS STORED, NOT La.  X[n] = Ls[n] + La[n] is kept per information bit; each SISO reads it
//    (SISO-2 through the QPP permutation), and overwrites it in place with Ls + its own scaled
//    extrinsic, which is exactly the other SISO's U.  Interleave/de-interleave (:54-96,
//    :1221,:1242) are thus the addressing of one read and one write, conflict-free in the
//    step-major layout (tdb200_internal.h).
//  * Tail bits only shape the beta vector at step K (La is zero there, :1224-1227), so they are
//    folded into a constant start vector once per decode.
//  * EARLY TERMINATION (new functionality; the reference only has a placeholder,
//    previous/Decoder.cc:1098-1099): every thread keeps the hard decisions of its own steps as a
//    bit mask; a codeblock stops after the first iteration (>= 2) that changes none of them and
//    leaves every |a-posteriori| at or above a threshold.  A CTA leaves when all its codeblocks did.
//  * THE ALU PIPE IS THE BOUND (VIADDMNMX issues there and nowhere else), so address arithmetic,
//    bit-wise NOT, byte packing and the 3/4 extrinsic scale are written as IMADs with opaque
//    constants (PassCfg) to keep them on the fma-heavy pipe; the backward sweep is ONE rolled
//    window body per pass (instruction-cache footprint).
//
// The kernel is a template over the geometry: K = 6144 / 5120 / 4096 (128 sub-blocks of 48 / 40 / 32
// steps, guard 16) get compile-time constants -- every shared-memory address becomes
// base+immediate -- and every other LTE size runs the same code with run-time geometry, several
// codeblock pairs side by side in one CTA when a codeblock needs fewer than 32 threads.
//
// This header holds the device code and the per-input-type kernel table; it is instantiated once
// per channel-LLR type in its own translation unit (tdb200_fast_inst_*.cu) so that the ~12 kernel
// variants of each type compile in parallel.  The host side is tdb200_fast.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

typedef uint32_t w32;  // two int16 lanes: codeblock A in bits 0-15, codeblock B in bits 16-31

__device__ __forceinline__ w32 vadd(w32 a, w32 b) { return __vadd2(a, b); }                        // VIADD.16x2      (fma-heavy pipe)
__device__ __forceinline__ w32 vaddmax(w32 a, w32 b, w32 c) { return __viaddmax_s16x2(a, b, c); }  // VIADDMNMX.S16x2 (alu pipe): max(a+b,c)
__device__ __forceinline__ w32 vneg(w32 a) { return __vadd2(~a, 0x00010001u); }
__device__ __forceinline__ w32 pack2(int lo, int hi) { return (w32)(lo & 0xffff) | ((w32)hi << 16); }
__device__ __forceinline__ w32 dup2(int v) { return pack2(v, v); }

__device__ __forceinline__ void norm8(w32 (&m)[8])
{
    const w32 nz = vneg(m[0]);
    m[0] = 0;
#pragma unroll
    for (int s = 1; s < 8; s++) m[s] = vadd(m[s], nz);
}

// alpha(i+1) from alpha(i); u = U_i, v = V_i   (the max-log form of :975-1001)
__device__ __forceinline__ void alpha_step_to(const w32 (&a)[8], w32 u, w32 v, w32 (&o)[8])
{
    const w32 w = vadd(u, v);
    const w32 t5 = vadd(a[2], v), t1 = vadd(a[3], v), t2 = vadd(a[4], v), t6 = vadd(a[5], v);
    const w32 o0 = vaddmax(a[1], w, a[0]), o4 = vaddmax(a[0], w, a[1]);
    const w32 o5 = vaddmax(a[3], u, t5), o1 = vaddmax(a[2], u, t1);
    const w32 o2 = vaddmax(a[5], u, t2), o6 = vaddmax(a[4], u, t6);
    const w32 o7 = vaddmax(a[7], w, a[6]), o3 = vaddmax(a[6], w, a[7]);
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5; o[6] = o6; o[7] = o7;
}
__device__ __forceinline__ void alpha_step(w32 (&a)[8], w32 u, w32 v) { alpha_step_to(a, u, v, a); }

// beta(i) from beta(i+1)   (:1004-1021)
__device__ __forceinline__ void beta_step(w32 (&b)[8], w32 u, w32 v)
{
    const w32 w = vadd(u, v);
    const w32 t2 = vadd(b[5], v), t3 = vadd(b[1], v), t4 = vadd(b[2], v), t5 = vadd(b[6], v);
    const w32 o0 = vaddmax(b[4], w, b[0]), o1 = vaddmax(b[0], w, b[4]);
    const w32 o2 = vaddmax(b[1], u, t2), o3 = vaddmax(b[5], u, t3);
    const w32 o4 = vaddmax(b[6], u, t4), o5 = vaddmax(b[2], u, t5);
    const w32 o6 = vaddmax(b[3], w, b[7]), o7 = vaddmax(b[7], w, b[3]);
    b[0] = o0; b[1] = o1; b[2] = o2; b[3] = o3; b[4] = o4; b[5] = o5; b[6] = o6; b[7] = o7;
}

// Constants the compiler must not see through: ptxas strength-reduces x*-1-1, x*2, x*65536 and
// mulhi(x, 3<<30) into ALU-pipe instructions (IADD3 / LEA / PRMT / SHF), and the ALU pipe is the one
// the add-compare-select instructions saturate.  Passed as kernel arguments they stay IMADs on the
// fma-heavy pipe, which has slack.
struct PassCfg {
    int q2;
    w32 lim1;    // dup2(ext_lim + 1): the +1 completes m1 + ~m0 = m1 - m0 - 1
    w32 limmax;  // dup2(2*ext_lim - 1)
    w32 unbias;  // dup2(-(3*ext_lim/4) - 128) or dup2(-ext_lim - 128): extrinsic bias and systematic-byte bias
    w32 neg1;    // 0xffffffff
    w32 four;    // 4
    w32 k64k;    // 65536
    w32 k3q;     // 0xC0000000: mulhi(y, k3q) = (3*y) >> 2
    w32 etT, et2T, etmask;  // dup2(T), dup2(2T), dup2(2T-1): magnitude test of the stopping rule
    // TDB200_ALGO_LOGMAP_S16 (see "max* with the correction" below)
    w32 lk;   // dup2(1 + T4 - 0x4000): completes every sum of two biased quarter-differences
    w32 lku;  // dup2(1 + T4): the a-posteriori trees' upper levels
    w32 lm, lx2, lx1, lkc;  // 0x3fff3fff, 0x20002000, 0x1fff1fff, 0xC000C000
};

// ---- max* with the correction: TDB200_ALGO_LOGMAP_S16 (integer specification: oracle/turbo_oracle_fx.c, logmap = 1)
//
// max*(x,y) = max(x,y) + c(|x-y|) is the Jacobian logarithm the reference evaluates through E_algorithm()'s
// 16-step table (ITTC/log_map.cpp:779-801, table :14-18; 30 of them per trellis step, :975-1039).  Here c is the
// linear fit max(0, T4 - |d|/4) in units of 2^-frac_bits (T4 = 10 at 4 fractional bits: 0.625 - d/4).
// sm_100 has no packed 16-bit shift, subtract or absolute difference, so d/4 is never formed per max*:
//   * the two max* of a trellis butterfly, max*(p+g, q) and max*(q+g, p), have d = (p-q)+g and d = g-(p-q).
//     One 32-bit SHF + two LOP3 per butterfly give h' = floor((p-q-1)/4) + 0x2000 and nh' = -floor(..)-1 + 0x2000 in
//     both lanes (the 14-bit field is re-biased, which undoes the cross-lane bits of the 32-bit shift), and the same
//     pair once per step for each of the two branch-metric differences g (u+v and u-v);
//   * T4 - |d/4| = min(T4 + d/4, T4 - d/4) is then ONE VIADDMNMX.RELU per max*: relu(min(h' + GP, nh' + GM));
//   * the first level of the a-posteriori trees pairs (alpha_i, alpha_j) x (beta_m, beta_n) along the same
//     butterflies, so it reuses the h' of the alpha and beta recursions; only the six upper-level max* per step pay
//     a shift each (via max - min, whose sign is known).
struct GK {
    w32 P, M;  // floor(g/4) + 1 + T4 - 0x2000,  T4 - floor(g/4) - 0x2000
};
#ifndef TDB_LM_NOT_FMA
#define TDB_LM_NOT_FMA 1
#endif
#ifndef TDB_LM_ALPHA_NOT_ALU
#define TDB_LM_ALPHA_NOT_ALU 1  // the alpha butterflies' NOTs as LOP3: evens out the two pipes (ncu: fma-heavy 74 %, ALU 65 % busy without)
#endif
// ~x: as an IMAD with opaque constants (fma-heavy pipe) or as a LOP3 (ALU pipe) -- whichever pipe has slack
__device__ __forceinline__ w32 vnot(w32 x, const PassCfg &c) { return TDB_LM_NOT_FMA ? x * c.neg1 + c.neg1 : ~x; }
__device__ __forceinline__ w32 and_xor(w32 a, w32 b, w32 x)  // (a & b) ^ x in one LOP3 (the compiler shares the AND and spends three)
{
    w32 r;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(r) : "r"(a), "r"(b), "r"(x));
    return r;
}
__device__ __forceinline__ void quarter(w32 d1, const PassCfg &c, w32 &h, w32 &nh)
{
    const w32 sh = d1 >> 2;
    h = and_xor(sh, c.lm, c.lx2);
    nh = and_xor(sh, c.lm, c.lx1);
}
__device__ __forceinline__ GK gk_of(w32 g, const PassCfg &c)
{
    w32 a, b;
    quarter(g, c, a, b);
    GK r;
    r.P = vadd(a, c.lk);
    r.M = vadd(b, c.lk);
    return r;
}
// corrections of max*(p + g, q) and max*(q + g, p), given the quarter-difference of (p, q)
__device__ __forceinline__ void bfly_corr(w32 h, w32 nh, const GK &g, w32 &ca, w32 &cb)
{
    ca = __viaddmin_s16x2_relu(h, g.P, vadd(nh, g.M));
    cb = __viaddmin_s16x2_relu(nh, g.P, vadd(h, g.M));
}
// quarter-differences of the four butterflies of an alpha vector: (a1,a0) (a3,a2) (a5,a4) (a7,a6)
__device__ __forceinline__ void alpha_quarters(const w32 (&a)[8], const PassCfg &c, w32 (&h)[4], w32 (&nh)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) quarter(vadd(a[2 * i + 1], TDB_LM_ALPHA_NOT_ALU ? ~a[2 * i] : vnot(a[2 * i], c)), c, h[i], nh[i]);
}
// ... of a beta vector: (b4,b0) (b1,b5) (b6,b2) (b3,b7)
__device__ __forceinline__ void beta_quarters(const w32 (&b)[8], const PassCfg &c, w32 (&h)[4], w32 (&nh)[4])
{
    quarter(vadd(b[4], vnot(b[0], c)), c, h[0], nh[0]);
    quarter(vadd(b[1], vnot(b[5], c)), c, h[1], nh[1]);
    quarter(vadd(b[6], vnot(b[2], c)), c, h[2], nh[2]);
    quarter(vadd(b[3], vnot(b[7], c)), c, h[3], nh[3]);
}
__device__ __forceinline__ void alpha_step_lm_to(const w32 (&a)[8], const w32 (&h)[4], const w32 (&nh)[4], w32 u, w32 v, const PassCfg &c,
                                                 w32 (&o)[8])
{
    const w32 w = vadd(u, v);
    const GK gw = gk_of(w, c), gg = gk_of(vadd(u, vnot(v, c)), c);
    w32 c0, c4, c5, c1, c2, c6, c7, c3;
    bfly_corr(h[0], nh[0], gw, c0, c4);
    bfly_corr(h[1], nh[1], gg, c5, c1);
    bfly_corr(h[2], nh[2], gg, c2, c6);
    bfly_corr(h[3], nh[3], gw, c7, c3);
    const w32 t5 = vadd(a[2], v), t1 = vadd(a[3], v), t2 = vadd(a[4], v), t6 = vadd(a[5], v);
    const w32 o0 = vadd(vaddmax(a[1], w, a[0]), c0), o4 = vadd(vaddmax(a[0], w, a[1]), c4);
    const w32 o5 = vadd(vaddmax(a[3], u, t5), c5), o1 = vadd(vaddmax(a[2], u, t1), c1);
    const w32 o2 = vadd(vaddmax(a[5], u, t2), c2), o6 = vadd(vaddmax(a[4], u, t6), c6);
    const w32 o7 = vadd(vaddmax(a[7], w, a[6]), c7), o3 = vadd(vaddmax(a[6], w, a[7]), c3);
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5; o[6] = o6; o[7] = o7;
}
__device__ __forceinline__ void beta_step_lm(w32 (&b)[8], const w32 (&h)[4], const w32 (&nh)[4], w32 u, w32 v, const PassCfg &c)
{
    const w32 w = vadd(u, v);
    const GK gw = gk_of(w, c), gg = gk_of(vadd(u, vnot(v, c)), c);
    w32 c0, c1, c2, c3, c4, c5, c6, c7;
    bfly_corr(h[0], nh[0], gw, c0, c1);
    bfly_corr(h[1], nh[1], gg, c2, c3);
    bfly_corr(h[2], nh[2], gg, c4, c5);
    bfly_corr(h[3], nh[3], gw, c6, c7);
    const w32 t2 = vadd(b[5], v), t3 = vadd(b[1], v), t4 = vadd(b[2], v), t5 = vadd(b[6], v);
    const w32 o0 = vadd(vaddmax(b[4], w, b[0]), c0), o1 = vadd(vaddmax(b[0], w, b[4]), c1);
    const w32 o2 = vadd(vaddmax(b[1], u, t2), c2), o3 = vadd(vaddmax(b[5], u, t3), c3);
    const w32 o4 = vadd(vaddmax(b[6], u, t4), c4), o5 = vadd(vaddmax(b[2], u, t5), c5);
    const w32 o6 = vadd(vaddmax(b[3], w, b[7]), c6), o7 = vadd(vaddmax(b[7], w, b[3]), c7);
    b[0] = o0; b[1] = o1; b[2] = o2; b[3] = o3; b[4] = o4; b[5] = o5; b[6] = o6; b[7] = o7;
}
// one recursion step, quarter-differences formed on the spot (warm-up, forward sweep, tail)
template <bool LM>
__device__ __forceinline__ void alpha_step_x(w32 (&a)[8], w32 u, w32 v, const PassCfg &c)
{
    if (LM) {
        w32 h[4], nh[4];
        alpha_quarters(a, c, h, nh);
        alpha_step_lm_to(a, h, nh, u, v, c, a);
    } else {
        alpha_step(a, u, v);
    }
}
template <bool LM>
__device__ __forceinline__ void beta_step_x(w32 (&b)[8], w32 u, w32 v, const PassCfg &c)
{
    if (LM) {
        w32 h[4], nh[4];
        beta_quarters(b, c, h, nh);
        beta_step_lm(b, h, nh, u, v, c);
    } else {
        beta_step(b, u, v);
    }
}
// max* of two values that share nothing (upper levels of the a-posteriori trees): mn - mx - 1 is negative in both
// lanes, so its arithmetic shift is the 32-bit shift with the two top bits of each lane set
__device__ __forceinline__ w32 maxstar_g(w32 x, w32 y, const PassCfg &c)
{
    const w32 mx = __vmaxs2(x, y), mn = __vmins2(x, y);
    const w32 e4 = (vadd(mn, vnot(mx, c)) >> 2) | c.lkc;
    return vadd(mx, __viaddmax_s16x2_relu(e4, c.lku, 0u));
}
// first level: the input-0 terms alpha_i + beta_m, alpha_j + beta_n ("same") and the input-1 terms alpha_j + beta_m,
// alpha_i + beta_n ("cross") of the state pair (i, j) x (m, n); hA / hB: quarter-differences of (aj, ai) / (bn, bm)
__device__ __forceinline__ void lam_pair(w32 ai, w32 aj, w32 bm, w32 bn, w32 hA, w32 nhA, w32 hB, w32 nhB, const PassCfg &c, w32 &same, w32 &cross)
{
    const w32 HB = vadd(hB, c.lk), nHB = vadd(nhB, c.lk);
    const w32 c1 = __viaddmin_s16x2_relu(hA, HB, vadd(nhA, nHB));
    const w32 c2 = __viaddmin_s16x2_relu(hA, nHB, vadd(nhA, HB));
    same = vadd(vaddmax(aj, bn, vadd(ai, bm)), c1);
    cross = vadd(vaddmax(aj, bm, vadd(ai, bn)), c2);
}
// e - 1 like extrinsic_m1 below, every max a max* (:1024-1039)
__device__ __forceinline__ w32 extrinsic_m1_lm(const w32 (&a)[8], const w32 (&hA)[4], const w32 (&b)[8], const w32 (&hB)[4], const w32 (&nhB)[4], w32 v,
                                               const PassCfg &c)
{
    w32 s01, x01, s67, x67, s23, x23, s45, x45;
    lam_pair(a[0], a[1], b[0], b[4], hA[0], hA[0] * c.neg1 + c.lm, hB[0], nhB[0], c, s01, x01);
    lam_pair(a[6], a[7], b[7], b[3], hA[3], hA[3] * c.neg1 + c.lm, hB[3], nhB[3], c, s67, x67);
    lam_pair(a[2], a[3], b[5], b[1], hA[1], hA[1] * c.neg1 + c.lm, hB[1], nhB[1], c, s23, x23);
    lam_pair(a[4], a[5], b[2], b[6], hA[2], hA[2] * c.neg1 + c.lm, hB[2], nhB[2], c, s45, x45);
    const w32 m0a = maxstar_g(s01, s67, c), m0b = maxstar_g(s23, s45, c);
    const w32 m1a = maxstar_g(x01, x67, c), m1b = maxstar_g(x23, x45, c);
    const w32 m0 = maxstar_g(m0a, vadd(m0b, v), c);
    const w32 m1 = maxstar_g(vadd(m1a, v), m1b, c);
    return vadd(m1, vnot(m0, c));
}

// ---- channel-LLR load + quantisation (q = clamp(rint(x * 2^F), +-clip), oracle: quant())
__device__ __forceinline__ int quant(float x, float scale, int clip)
{
    float s = x * scale;
    if (!(s == s)) return 0;
    s = fminf(fmaxf(s, -32767.0f), 32767.0f);
    const int q = __float2int_rn(s);
    return max(min(q, clip), -clip);
}

// Two channel values (codeblock A, codeblock B) -> one clipped s16x2.  cvt.rni.sat.s16.f32 rounds
// to nearest even, saturates and maps NaN to 0, exactly like quant() above for |clip| <= 32767.
__device__ __forceinline__ w32 quant2(float a, float b, float scale, w32 clipv, w32 nclipv)
{
    short qa, qb;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(qa) : "f"(a * scale));
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(qb) : "f"(b * scale));
    w32 p;
    asm("mov.b32 %0, {%1, %2};" : "=r"(p) : "h"(qa), "h"(qb));
    return __vmins2(__vmaxs2(p, nclipv), clipv);
}

// 12 consecutive input values (4 systematic/parity1/parity2 triplets) of one codeblock, starting at
// element 12*q, as loaded (128-bit loads for the float types)
template <int LLR_T>
struct Raw12 {
    float4 f[(LLR_T == TDB200_LLR_F32 || LLR_T == kLlrSymBpskF32) ? 3 : 1];
    double2 d[LLR_T == TDB200_LLR_F64 ? 6 : 1];
    int w[LLR_T == TDB200_LLR_S8 ? 3 : 1];
    uint2 h[LLR_T == TDB200_LLR_F16 ? 3 : 1];  // 4 halves each
    float2 si[LLR_T == kLlrSymQpskF32 ? 3 : 1], sq[LLR_T == kLlrSymQpskF32 ? 3 : 1];  // six QPSK symbols: I and Q planes
};

// The soft demapper fused into the load stage (BPSK and QPSK, float symbols): the max-log metric of demodule()
// (ITTC/modanddem.cpp:189-260) exactly as demap32_kernel evaluates it per axis (csrc/tdb200_modem.cu: axis32<1>) --
// LLR = -Kf ((v - l1)^2 - (v - l0)^2) with the axis' two levels (l0, l1) = (-1, 1) for BPSK, (0.7071, -0.7071) for QPSK,
// every operation rounded on its own (no contraction), so that decoding symbols equals demapping to the 8-bit
// hand-over format and decoding that.
template <int LLR_T>
__device__ __forceinline__ float demap1(float v, float kf)
{
    const float l0 = LLR_T == kLlrSymBpskF32 ? -1.0f : 0.7071f, l1 = LLR_T == kLlrSymBpskF32 ? 1.0f : -0.7071f;
    const float t1 = __fsub_rn(v, l1), t0 = __fsub_rn(v, l0);
    return __fmul_rn(-kf, __fsub_rn(__fmul_rn(t1, t1), __fmul_rn(t0, t0)));
}

template <int LLR_T>
__device__ __forceinline__ void load12(const FastArgs &A, size_t row_elems, int cb, int q, Raw12<LLR_T> &r)
{
    const void *base = A.llr;
    if (LLR_T == kLlrSymQpskF32) {
        // 12 soft bits = 6 symbols; rows hold (3K+12)/2 symbols per plane (an even number: 8-byte aligned rows)
        const float2 *pi = reinterpret_cast<const float2 *>(static_cast<const float *>(A.llr) + (size_t)cb * (row_elems / 2)) + 3 * q;
        const float2 *pq = reinterpret_cast<const float2 *>(static_cast<const float *>(A.sym_q) + (size_t)cb * (row_elems / 2)) + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) { r.si[k] = __ldg(pi + k); r.sq[k] = __ldg(pq + k); }
    } else if (LLR_T == TDB200_LLR_F32 || LLR_T == kLlrSymBpskF32) {
        const float4 *p = reinterpret_cast<const float4 *>(static_cast<const float *>(base) + (size_t)cb * row_elems) + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) r.f[k] = __ldg(p + k);
    } else if (LLR_T == TDB200_LLR_F64) {
        const double2 *p = reinterpret_cast<const double2 *>(static_cast<const double *>(base) + (size_t)cb * row_elems) + 6 * q;
#pragma unroll
        for (int k = 0; k < 6; k++) r.d[k] = __ldg(p + k);
    } else if (LLR_T == TDB200_LLR_F16) {
        const uint2 *p = reinterpret_cast<const uint2 *>(static_cast<const __half *>(base) + (size_t)cb * row_elems) + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) r.h[k] = __ldg(p + k);
    } else {
        const int *p = reinterpret_cast<const int *>(static_cast<const int8_t *>(base) + (size_t)cb * row_elems) + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) r.w[k] = __ldg(p + k);
    }
}
__device__ __forceinline__ float h_lo(unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float h_hi(unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

// quantise + pack the 12 values of codeblocks A and B into 12 s16x2 words
template <int LLR_T>
__device__ __forceinline__ void pack12(const Raw12<LLR_T> &a, const Raw12<LLR_T> &b, float scale, w32 clipv, w32 nclipv, w32 (&out)[12], float kf)
{
    if (LLR_T == kLlrSymBpskF32) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            out[4 * k] = quant2(demap1<LLR_T>(a.f[k].x, kf), demap1<LLR_T>(b.f[k].x, kf), scale, clipv, nclipv);
            out[4 * k + 1] = quant2(demap1<LLR_T>(a.f[k].y, kf), demap1<LLR_T>(b.f[k].y, kf), scale, clipv, nclipv);
            out[4 * k + 2] = quant2(demap1<LLR_T>(a.f[k].z, kf), demap1<LLR_T>(b.f[k].z, kf), scale, clipv, nclipv);
            out[4 * k + 3] = quant2(demap1<LLR_T>(a.f[k].w, kf), demap1<LLR_T>(b.f[k].w, kf), scale, clipv, nclipv);
        }
    } else if (LLR_T == kLlrSymQpskF32) {
#pragma unroll
        for (int k = 0; k < 3; k++) {   // soft bit 2m from I[m], 2m+1 from Q[m]
            out[4 * k] = quant2(demap1<LLR_T>(a.si[k].x, kf), demap1<LLR_T>(b.si[k].x, kf), scale, clipv, nclipv);
            out[4 * k + 1] = quant2(demap1<LLR_T>(a.sq[k].x, kf), demap1<LLR_T>(b.sq[k].x, kf), scale, clipv, nclipv);
            out[4 * k + 2] = quant2(demap1<LLR_T>(a.si[k].y, kf), demap1<LLR_T>(b.si[k].y, kf), scale, clipv, nclipv);
            out[4 * k + 3] = quant2(demap1<LLR_T>(a.sq[k].y, kf), demap1<LLR_T>(b.sq[k].y, kf), scale, clipv, nclipv);
        }
    } else if (LLR_T == TDB200_LLR_F32) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            out[4 * k] = quant2(a.f[k].x, b.f[k].x, scale, clipv, nclipv);
            out[4 * k + 1] = quant2(a.f[k].y, b.f[k].y, scale, clipv, nclipv);
            out[4 * k + 2] = quant2(a.f[k].z, b.f[k].z, scale, clipv, nclipv);
            out[4 * k + 3] = quant2(a.f[k].w, b.f[k].w, scale, clipv, nclipv);
        }
    } else if (LLR_T == TDB200_LLR_F64) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            out[2 * k] = quant2((float)a.d[k].x, (float)b.d[k].x, scale, clipv, nclipv);
            out[2 * k + 1] = quant2((float)a.d[k].y, (float)b.d[k].y, scale, clipv, nclipv);
        }
    } else if (LLR_T == TDB200_LLR_F16) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            out[4 * k] = quant2(h_lo(a.h[k].x), h_lo(b.h[k].x), scale, clipv, nclipv);
            out[4 * k + 1] = quant2(h_hi(a.h[k].x), h_hi(b.h[k].x), scale, clipv, nclipv);
            out[4 * k + 2] = quant2(h_lo(a.h[k].y), h_lo(b.h[k].y), scale, clipv, nclipv);
            out[4 * k + 3] = quant2(h_hi(a.h[k].y), h_hi(b.h[k].y), scale, clipv, nclipv);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            // bytes {a_m, sign(a_m), b_m, sign(b_m)}; prmt, not __byte_perm: the intrinsic drops the
            // sign-replication bit of the selector
            w32 p0, p1, p2, p3;
            asm("prmt.b32 %0, %1, %2, 0xc480;" : "=r"(p0) : "r"(a.w[k]), "r"(b.w[k]));
            asm("prmt.b32 %0, %1, %2, 0xd591;" : "=r"(p1) : "r"(a.w[k]), "r"(b.w[k]));
            asm("prmt.b32 %0, %1, %2, 0xe6a2;" : "=r"(p2) : "r"(a.w[k]), "r"(b.w[k]));
            asm("prmt.b32 %0, %1, %2, 0xf7b3;" : "=r"(p3) : "r"(a.w[k]), "r"(b.w[k]));
            out[4 * k] = __vmins2(__vmaxs2(p0, nclipv), clipv);
            out[4 * k + 1] = __vmins2(__vmaxs2(p1, nclipv), clipv);
            out[4 * k + 2] = __vmins2(__vmaxs2(p2, nclipv), clipv);
            out[4 * k + 3] = __vmins2(__vmaxs2(p3, nclipv), clipv);
        }
    }
}

template <int LLR_T>
__device__ __forceinline__ int load1(const FastArgs &A, int cb, size_t row_elems, size_t o, float scale, int clip)
{
    const void *base = A.llr;
    const size_t idx = (size_t)cb * row_elems + o;
    if (LLR_T == kLlrSymBpskF32) return quant(demap1<LLR_T>(__ldg(static_cast<const float *>(base) + idx), A.kf), scale, clip);
    if (LLR_T == kLlrSymQpskF32) {
        const float *pl = static_cast<const float *>((o & 1) ? A.sym_q : A.llr);
        return quant(demap1<LLR_T>(__ldg(pl + (size_t)cb * (row_elems / 2) + (o >> 1)), A.kf), scale, clip);
    }
    if (LLR_T == TDB200_LLR_F32) return quant(__ldg(static_cast<const float *>(base) + idx), scale, clip);
    if (LLR_T == TDB200_LLR_F64) return quant((float)__ldg(static_cast<const double *>(base) + idx), scale, clip);
    if (LLR_T == TDB200_LLR_F16) return quant(__half2float(__ldg(static_cast<const __half *>(base) + idx)), scale, clip);
    const int v = (int)__ldg(static_cast<const int8_t *>(base) + idx);
    return max(min(v, clip), -clip);
}

struct Smem {
    w32 *X, *par1, *par2;
    uint8_t *sysA, *sysB;  // systematic value + 128 of codeblock A / B, one byte per trellis step
    uint16_t *tab;
    w32 *ckpt, *dec, *edge;
    w32 *tail;  // [2][8]: the two SISOs' beta vectors at step K (tail bits folded in), for warm-ups that reach the trellis end
};

__device__ __forceinline__ w32 vnot_fma(w32 x, w32 neg1) { return x * neg1 + neg1; }  // ~x

#ifndef TDB_LAMBDA_V3
#define TDB_LAMBDA_V3 2
#endif

// max of four (alpha + beta') sums.  Form A: 1 fma-pipe + 3 ALU-pipe instructions; form B (three-input
// maximum): 3 fma-pipe + 2 ALU-pipe.  TDB_LAMBDA_V3 of the four groups per step use form B, which is
// what balances the two pipes in the backward window.
template <bool V3>
__device__ __forceinline__ w32 max4sum(w32 a0, w32 b0, w32 a1, w32 b1, w32 a2, w32 b2, w32 a3, w32 b3)
{
    if (V3) return vaddmax(a3, b3, __vimax3_s16x2(vadd(a0, b0), vadd(a1, b1), vadd(a2, b2)));
    return vaddmax(a3, b3, vaddmax(a2, b2, vaddmax(a1, b1, vadd(a0, b0))));
}

// e - 1, where e = max_{input 1}(alpha + c*V + beta') - max_{input 0}(alpha + c*V + beta')
// (:1024-1039 as max-log; the +U common to all input-1 branches is left out, so e IS the extrinsic
// of :1234-1238).  The -1 is absorbed by the clamp constant.
__device__ __forceinline__ w32 extrinsic_m1(const w32 (&a)[8], const w32 (&b)[8], w32 v, w32 neg1)
{
    const w32 m0a = max4sum<(TDB_LAMBDA_V3 > 0)>(a[0], b[0], a[1], b[4], a[6], b[7], a[7], b[3]);
    const w32 m0b = max4sum<(TDB_LAMBDA_V3 > 2)>(a[2], b[5], a[3], b[1], a[4], b[2], a[5], b[6]);
    const w32 m1a = max4sum<(TDB_LAMBDA_V3 > 3)>(a[0], b[4], a[1], b[0], a[6], b[3], a[7], b[7]);
    const w32 m1b = max4sum<(TDB_LAMBDA_V3 > 1)>(a[2], b[1], a[3], b[5], a[4], b[6], a[5], b[2]);
    const w32 m0 = vaddmax(m0b, v, m0a);
    const w32 m1 = vaddmax(m1a, v, m1b);
    return vadd(m1, vnot_fma(m0, neg1));
}

// Where a trellis step's a-priori word X and systematic byte pair live.  Natural-order passes
// address them by step index (base + immediate); interleaved passes go through tab, which holds the
// word index e of the element: its byte offset in the systematic planes, a quarter of the one in X.
struct Elem {
    unsigned xoff, soff;  // byte offsets into X / the systematic byte planes
};
__device__ __forceinline__ Elem elem_of(const bool IL, const PassCfg &c, unsigned tabval, int idx)
{
    Elem e;
    e.soff = IL ? tabval : (unsigned)idx;
    e.xoff = IL ? tabval * c.four : 4u * (unsigned)idx;
    return e;
}
__device__ __forceinline__ Elem elem_at(const bool IL, const PassCfg &c, const Smem &sm, int idx)
{
    return elem_of(IL, c, IL ? (unsigned)sm.tab[idx] : 0u, idx);
}
__device__ __forceinline__ w32 &word_at(w32 *base, unsigned xoff)
{
    return *reinterpret_cast<w32 *>(reinterpret_cast<unsigned char *>(base) + xoff);
}
__device__ __forceinline__ w32 &x_at(const Smem &sm, const Elem &e) { return word_at(sm.X, e.xoff); }
// systematic values of the two codeblocks, each + 128, as an s16x2: one byte load per codeblock
// (separate planes, so the compiler cannot merge them into a 16-bit load + two PRMTs) and one IMAD
__device__ __forceinline__ w32 sys_biased(const PassCfg &c, const Smem &sm, const Elem &e)
{
    return (w32)sm.sysB[e.soff] * c.k64k + (w32)sm.sysA[e.soff];
}

// Backward sweep over one 8-step window whose first alpha vector is a0 (normalised): re-create
// the window's alpha vectors in registers, then run beta, the extrinsic output and the in-place
// update of X over it.  tabin: the window's table entries (interleaved passes fetch them one window
// ahead, so the look-up is off the critical path).  Returns the decision bits of the window (WANT
// only): sign of step k in bit 15-k (codeblock A) / 31-k (codeblock B); weak collects, per lane, a
// non-zero value if some |a-posteriori| of the window is below the stopping threshold.
template <bool LM>
__device__ __forceinline__ w32 bwd_window(const bool IL, const bool WANT, const PassCfg &c, const Smem &sm, const w32 *par, const int base, const int PP,
                                          const unsigned (&tabin)[8], const w32 (&a0)[8], w32 (&b)[8], w32 *stage, w32 &weak)
{
    w32 aw[8][8], u[8], v[8];
    w32 hA[LM ? 8 : 1][4];  // Log-MAP: quarter-differences of the window's alpha vectors, shared by the re-creation and the a-posteriori trees
    Elem el[8];
#pragma unroll
    for (int s = 0; s < 8; s++) aw[0][s] = a0[s];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int idx = base + k * PP;
        el[k] = elem_of(IL, c, tabin[k], idx);
        u[k] = x_at(sm, el[k]);
        v[k] = par[idx];
        if (LM) {
            w32 nh[4];
            alpha_quarters(aw[k], c, hA[k], nh);
            if (k < 7) alpha_step_lm_to(aw[k], hA[k], nh, u[k], v[k], c, aw[k + 1]);
        } else {
            if (k < 7) alpha_step_to(aw[k], u[k], v[k], aw[k + 1]);
        }
    }
    norm8(b);
    w32 acc = 0;
    if (LM) {
        // The beta / a-posteriori body of a Log-MAP step is ~210 instructions: unrolled eight times the window would
        // be 37 KB of code, more than the 32 KB instruction cache in front of the SM.  It runs as a ROLLED loop over
        // the two halves of the window instead; the working registers of the second half are moved into place.
        w32 wa[4][8], wh[4][4], wu[4], wv[4];
        Elem we[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int s = 0; s < 8; s++) wa[k][s] = aw[4 + k][s];
#pragma unroll
            for (int s = 0; s < 4; s++) wh[k][s] = hA[4 + k][s];
            wu[k] = u[4 + k]; wv[k] = v[4 + k]; we[k] = el[4 + k];
        }
#pragma unroll 1
        for (int half = 1; half >= 0; half--) {
#pragma unroll
            for (int k = 3; k >= 0; k--) {
                w32 hB[4], nhB[4];
                beta_quarters(b, c, hB, nhB);
                const w32 exm1 = extrinsic_m1_lm(wa[k], wh[k], b, hB, nhB, wv[k], c);
                const w32 y = __viaddmin_s16x2_relu(exm1, c.lim1, c.limmax);
                w32 es;
                if (c.q2 == 3) es = __umulhi(y, c.k3q) & 0x3fff3fffu;
                else es = y;
                x_at(sm, we[k]) = vadd(vadd(sys_biased(c, sm, we[k]), es), c.unbias);
                if (WANT) {
                    const w32 lam = vadd(vadd(wu[k], exm1), 0x00010001u);
                    acc = (acc >> 1) | (lam & 0x80008000u);
                    weak |= __viaddmin_s16x2_relu(lam, c.etT, c.et2T) & c.etmask;
                    if (stage) word_at(stage, we[k].xoff) = lam;
                }
                beta_step_lm(b, hB, nhB, wu[k], wv[k], c);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int s = 0; s < 8; s++) wa[k][s] = aw[k][s];
#pragma unroll
                for (int s = 0; s < 4; s++) wh[k][s] = hA[k][s];
                wu[k] = u[k]; wv[k] = v[k]; we[k] = el[k];
            }
        }
        return acc;
    }
#pragma unroll
    for (int k = 7; k >= 0; k--) {
        w32 hB[4], nhB[4];
        w32 exm1;
        if (LM) {
            beta_quarters(b, c, hB, nhB);
            exm1 = extrinsic_m1_lm(aw[k], hA[k], b, hB, nhB, v[k], c);
        } else {
            exm1 = extrinsic_m1(aw[k], b, v[k], c.neg1);
        }
        // clamp to [-lim, lim-1], bias to [0, 2lim-1]
        const w32 y = __viaddmin_s16x2_relu(exm1, c.lim1, c.limmax);
        w32 es;
        if (c.q2 == 3) es = __umulhi(y, c.k3q) & 0x3fff3fffu;  // floor(3(ec+lim)/4) per lane
        else es = y;
        x_at(sm, el[k]) = vadd(vadd(sys_biased(c, sm, el[k]), es), c.unbias);
        if (WANT) {
            const w32 lam = vadd(vadd(u[k], exm1), 0x00010001u);  // a-posteriori, :1038 (+ the dropped U)
            acc = (acc >> 1) | (lam & 0x80008000u);
            // clamp(lam + T, 0, 2T) is 0 or 2T exactly when |lam| >= T (lam >= T or lam <= -T)
            weak |= __viaddmin_s16x2_relu(lam, c.etT, c.et2T) & c.etmask;
            if (stage) word_at(stage, el[k].xoff) = lam;
        }
        if (LM) beta_step_lm(b, hB, nhB, u[k], v[k], c);
        else beta_step(b, u[k], v[k]);
    }
    return acc;
}

// One SISO pass of one sub-block (thread `t` of its codeblock pair; inactive threads only take part in
// the barriers and the boundary exchange).  IL = false: SISO-1 (natural order), true: SISO-2 (through tab).
// na/nb: boundary vectors (alpha G steps before the sub-block, beta G steps after it); on return
// they hold the vectors for the next iteration of this SISO.  With WANT the hard decisions of this
// pass go to sm.dec (one word per two windows) and the return value has bits 0-15 / 16-31 set
// where a decision of codeblock A / B differs from what sm.dec held before; weak gets a non-zero
// low / high half if some a-posteriori magnitude of codeblock A / B is below the stopping threshold.
template <int ILT, int WANTT, int KP, int KNW, int KG, bool LM>
__device__ __forceinline__ w32 siso_pass(const PassCfg &c, const FastGeom &g, const Smem &sm, const w32 *par, w32 (&na)[8], w32 (&nb)[8],
                                         const int t, const bool active, const bool first_fixed, const bool last_fixed, w32 *stage,
                                         w32 &weak, const w32 live = 0xffffffffu, const bool il_rt = false, const bool want_rt = false)
{
    // (which SISO this is -- the tail vector a long guard may need -- follows from the parity array: par2 <=> SISO-2)
    // ILT / WANTT: 0 / 1 compile-time, -1 run-time (the Log-MAP kernels keep ONE copy of the pass: its body is three
    // times the max-log one, and four inlined copies would be a quarter of a megabyte of code)
    const bool IL = ILT < 0 ? il_rt : (ILT != 0), WANT = WANTT < 0 ? want_rt : (WANTT != 0);
    const int P = KP > 0 ? KP : g.P, PP = KP > 0 ? (KP | 1) : g.PP, NW = KP ? KNW : g.NW, G = KP ? KG : g.G;  // KP < 0: run-time P
    const int L = 8 * NW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // A guard longer than the sub-block (G = 2L: the 8-step sub-blocks of K = 8 x prime, run-time geometry only) spans
    // D = 2 sub-blocks: thread t warms up from the START vector thread t-D had in the previous iteration.
    const int D = (KP == 0 && G > L) ? G / L : 1;
    const int w_sa = (L > G ? L - G : 0) >> 3, w_sb = G >> 3;
    w32 a[8], b[8], a0[8], sa[8], sb[8];
    w32 changed = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) { a[s] = na[s]; b[s] = nb[s]; sa[s] = 0; sb[s] = 0; }

    unsigned offn[8];  // interleaved passes: table entries of the window that is processed next
    if (active) {
#pragma unroll
        for (int k = 0; k < 8; k++) offn[k] = IL ? (unsigned)sm.tab[k * PP + t] : 0u;
        // ---- warm-up: alpha over the last G steps of sub-block t-1 and beta over the first G steps
        //      of sub-block t+1, advanced together (two independent dependency chains).  The two
        //      edge threads run it on their own sub-block and throw the result away, which keeps the
        //      loop free of divergence.
        const int ta = first_fixed ? t : t - 1, tb = last_fixed ? t : t + 1;
#pragma unroll 1
        for (int g0 = 0; g0 < G; g0 += 8) {
            int base_a = (L - G + g0) * PP + ta;
            int base_b = (G - 8 - g0) * PP + tb;
            if (KP == 0 && G > L) {
                // the eight steps of this chunk by trellis position: sub-block ua (ub), local step la (lb).  A chunk
                // that would lie outside the trellis runs on the thread's own data and is thrown away: where the walk
                // reaches the first (last) trellis step it restarts from the known vector.
                const int pa = t * L - G + g0, pb = (t + 1) * L + G - 8 - g0, K = P * L;
                int ua = pa / L, la = pa - ua * L, ub = pb / L, lb = pb - ub * L;
                if (pa < 0) { ua = t; la = 0; }
                if (pb + 8 > K) { ub = t; lb = 0; }
                base_a = la * PP + ua;
                base_b = lb * PP + ub;
                if (pa == 0) {
#pragma unroll
                    for (int s = 0; s < 8; s++) a[s] = s ? dup2(kFxNeg) : 0u;
                }
                if (pb + 8 == K) {
                    const w32 *tv = sm.tail + (par == sm.par2 ? 8 : 0);
#pragma unroll
                    for (int s = 0; s < 8; s++) b[s] = tv[s];
                }
            }
            norm8(a);
            norm8(b);
            if (LM) {  // rolled by four steps: instruction-cache footprint (see bwd_window)
#pragma unroll 1
                for (int k0 = 0; k0 < 8; k0 += 4) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int ia = base_a + (k0 + k) * PP, ib = base_b + (7 - k0 - k) * PP;
                        alpha_step_x<LM>(a, x_at(sm, elem_at(IL, c, sm, ia)), par[ia], c);
                        beta_step_x<LM>(b, x_at(sm, elem_at(IL, c, sm, ib)), par[ib], c);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int ia = base_a + k * PP, ib = base_b + (7 - k) * PP;
                    alpha_step_x<LM>(a, x_at(sm, elem_at(IL, c, sm, ia)), par[ia], c);
                    beta_step_x<LM>(b, x_at(sm, elem_at(IL, c, sm, ib)), par[ib], c);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if (first_fixed) a[s] = na[s];
            if (last_fixed) b[s] = nb[s];
        }
        if (G >= L) {
#pragma unroll
            for (int s = 0; s < 8; s++) sb[s] = b[s];
        }
        norm8(a);
#pragma unroll
        for (int s = 0; s < 8; s++) a0[s] = a[s];
        // ---- forward sweep over windows 0..NW-2, leaving a (normalised) checkpoint at the start of
        //      windows 1..NW-2; the start of window NW-1 stays in registers
#pragma unroll 1
        for (int w = 0; w < NW - 1; w++) {
            Elem el[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                el[k] = elem_of(IL, c, offn[k], (8 * w + k) * PP + t);
                if (IL) offn[k] = sm.tab[(8 * (w + 1) + k) * PP + t];
            }
            if (w > 0) {
                norm8(a);
#pragma unroll
                for (int s = 1; s < 8; s++) sm.ckpt[((w - 1) * 7 + (s - 1)) * P + t] = a[s];
            }
            const int base = 8 * w * PP + t;
#pragma unroll
            for (int k = 0; k < 8; k++) alpha_step_x<LM>(a, x_at(sm, el[k]), par[base + k * PP], c);
        }
        if (NW > 1) norm8(a);
    }
    __syncthreads();  // every warm-up read of X precedes every in-place update below
    if (active) {
        // alpha at local step L-G, for the right-hand neighbour's next iteration: the start vector
        // of window w_sa -- live in registers for the first and last window, a checkpoint otherwise
        if (G > 0) {
            if (w_sa == NW - 1) {
#pragma unroll
                for (int s = 0; s < 8; s++) sa[s] = a[s];
            } else if (w_sa == 0) {
#pragma unroll
                for (int s = 0; s < 8; s++) sa[s] = a0[s];
            }
        } else {  // alpha at the sub-block end
            w32 tmp[8];
#pragma unroll
            for (int s = 0; s < 8; s++) tmp[s] = a[s];
            const int base = 8 * (NW - 1) * PP + t;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int idx = base + k * PP;
                alpha_step_x<LM>(tmp, x_at(sm, elem_at(IL, c, sm, idx)), par[idx], c);
            }
#pragma unroll
            for (int s = 0; s < 8; s++) sa[s] = tmp[s];
        }
        // ---- backward sweep, ONE window body for all windows (instruction-cache footprint): the
        //      start vector of a window comes from `spec` for the last and the first window (the
        //      forward sweep's registers, then the saved a0) and from the checkpoints otherwise.
        //      Every start vector is normalised, so component 0 is always 0.
        w32 spec[8];
#pragma unroll
        for (int s = 0; s < 8; s++) spec[s] = a[s];
        w32 hold = 0;  // decision bits of an odd window waiting for its even partner
#pragma unroll 1
        for (int w = NW - 1; w >= 0; w--) {
            const bool mid = (w > 0) && (w < NW - 1);
            w32 aw0[8];
            aw0[0] = 0;
#pragma unroll
            for (int s = 1; s < 8; s++) {
                aw0[s] = spec[s];
                if (mid) aw0[s] = sm.ckpt[((w - 1) * 7 + (s - 1)) * P + t];
            }
            unsigned off[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                off[k] = offn[k];
                if (IL) offn[k] = sm.tab[(8 * max(w - 1, 0) + k) * PP + t];
            }
            const w32 acc = bwd_window<LM>(IL, WANT, c, sm, par, 8 * w * PP + t, PP, off, aw0, b, stage, weak);
            if (w == w_sb) {
#pragma unroll
                for (int s = 0; s < 8; s++) sb[s] = b[s];
            }
            if (WANT) {
                if (w & 1) hold = acc;
                else {
                    const w32 word = (acc >> 8) | hold;
                    const w32 old = sm.dec[(w >> 1) * P + t];
                    changed |= word ^ old;
                    // a codeblock that has met its stopping rule keeps the decisions it stopped with (live: 0xffff per
                    // lane still running), whatever its lane mate and the other pairs of the CTA go on to do
                    sm.dec[(w >> 1) * P + t] = (word & live) | (old & ~live);
                    hold = 0;
                }
            }
#pragma unroll
            for (int s = 0; s < 8; s++) spec[s] = a0[s];
        }
        if (G > 0 && w_sa > 0 && w_sa < NW - 1) {
#pragma unroll
            for (int s = 1; s < 8; s++) sa[s] = sm.ckpt[((w_sa - 1) * 7 + (s - 1)) * P + t];
        }
        norm8(sa);
        norm8(sb);
    }
    // ---- hand the boundary vectors to the neighbours (they use them in the next iteration):
    //      warp shuffles inside a warp, one shared-memory word per state across warp edges
    w32 up[8], dn[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        up[s] = __shfl_up_sync(0xffffffffu, sa[s], D);
        dn[s] = __shfl_down_sync(0xffffffffu, sb[s], D);
    }
    // edge words: [state][distance slot 0..1][warp]
    if (lane >= 32 - D) {
#pragma unroll
        for (int s = 0; s < 8; s++) sm.edge[(s * 2 + (lane - (32 - D))) * nwarps + warp] = sa[s];
    }
    if (lane < D) {
#pragma unroll
        for (int s = 0; s < 8; s++) sm.edge[((8 + s) * 2 + lane) * nwarps + warp] = sb[s];
    }
    __syncthreads();  // also orders this pass's X updates before the next pass's reads
    if (lane < D && warp > 0) {
#pragma unroll
        for (int s = 0; s < 8; s++) up[s] = sm.edge[(s * 2 + lane) * nwarps + warp - 1];
    }
    if (lane >= 32 - D && warp + 1 < nwarps) {
#pragma unroll
        for (int s = 0; s < 8; s++) dn[s] = sm.edge[((8 + s) * 2 + (lane + D - 32)) * nwarps + warp + 1];
    }
#pragma unroll
    for (int s = 0; s < 8; s++) {
        if (!first_fixed) na[s] = up[s];
        if (!last_fixed) nb[s] = dn[s];
    }
    return changed;
}

// ---- de-multiplex one group of four trellis steps (12 packed words: sys,par1,par2 x 4) into shared memory
__device__ __forceinline__ void put4(const Smem &sm, int q, int L, int PP, const w32 (&v)[12])
{
    const int n = 4 * q;  // L is a multiple of 8, so the four steps share their sub-block
    const int tt = n / L, j = n - tt * L;
    const int ad = j * PP + tt;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        sm.X[ad + m * PP] = v[3 * m];
        const w32 sb = vadd(v[3 * m], 0x00800080u);  // value + 128 in each lane
        sm.sysA[ad + m * PP] = (uint8_t)sb;
        sm.sysB[ad + m * PP] = (uint8_t)(sb >> 16);
        sm.par1[ad + m * PP] = v[3 * m + 1];
        sm.par2[ad + m * PP] = v[3 * m + 2];
    }
}

// Shared memory of one CTA: NP codeblock-pair regions (X, par1, par2, sysA, sysB, ckpt, dec), then the
// QPP table (one copy for all pairs), the warp-edge exchange words and the per-pair stop flags.
__device__ __forceinline__ Smem pair_smem(unsigned char *raw, const FastGeom &g, int P, int NW, int Wp, int NP, int p)
{
    Smem sm;
    unsigned char *r = raw + (size_t)p * g.pair_bytes;
    sm.X = reinterpret_cast<w32 *>(r);
    sm.par1 = sm.X + Wp;
    sm.par2 = sm.par1 + Wp;
    sm.sysA = reinterpret_cast<uint8_t *>(sm.par2 + Wp);
    sm.sysB = sm.sysA + Wp;
    sm.ckpt = reinterpret_cast<w32 *>(sm.sysB + Wp);
    sm.dec = sm.ckpt + (size_t)g.n_ckpt * 7 * P;
    sm.tail = sm.dec + (size_t)((NW + 1) / 2) * P;
    unsigned char *sh = raw + (((size_t)NP * g.pair_bytes + 15) & ~(size_t)15);  // pair regions are word-aligned only (bank-staggered)
    sm.tab = reinterpret_cast<uint16_t *>(sh);
    sm.edge = reinterpret_cast<w32 *>(sm.tab + Wp);
    return sm;
}

// a * b mod g over GF(2), 24-bit operands (g = x^24 + poly)
__device__ __forceinline__ unsigned crc_mulmod24(unsigned a, unsigned b, unsigned poly)
{
    unsigned r = 0;
#pragma unroll 1
    for (int i = 0; i < 24; i++) {
        r ^= (b & 1u) ? a : 0u;
        b >>= 1;
        a = ((a << 1) & 0xffffffu) ^ ((a & 0x800000u) ? poly : 0u);
    }
    return r;
}

// CRC stopping rule: remainders of codeblocks A and B over the natural-order decisions SISO-1 left in
// sm.dec.  Every thread divides the 8*NW decisions of its own sub-block byte by byte (table look-up),
// weights the remainder with x^(bits that follow) mod g, and the CTA XORs the 24-bit results together:
// a CRC is linear over GF(2).  Returns {remainder A, remainder B}, identical in all threads.
__device__ __forceinline__ uint2 crc_of_decisions(const FastArgs &A, const Smem &sm, unsigned *flags, int P, int NW, int t, bool active)
{
    unsigned ca = 0, cb = 0;
    if (active) {
        for (int w = 0; w < NW; w++) {
            const w32 word = sm.dec[(w >> 1) * P + t] >> (8 * (w & 1));  // sign bits: set = decision 0
            const unsigned da = (~word) & 0xffu, db = (~(word >> 16)) & 0xffu;
            ca = ((ca << 8) & 0xffffffu) ^ __ldg(A.crc_tab + (((ca >> 16) ^ da) & 0xffu));
            cb = ((cb << 8) & 0xffffffu) ^ __ldg(A.crc_tab + (((cb >> 16) ^ db) & 0xffu));
        }
        const unsigned m = __ldg(A.crc_shift + t);
        ca = crc_mulmod24(ca, m, A.crc_poly);
        cb = crc_mulmod24(cb, m, A.crc_poly);
    }
    ca = __reduce_xor_sync(0xffffffffu, ca);
    cb = __reduce_xor_sync(0xffffffffu, cb);
    if (threadIdx.x < 2) flags[threadIdx.x] = 0u;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        atomicXor(&flags[0], ca);
        atomicXor(&flags[1], cb);
    }
    __syncthreads();
    const uint2 r = make_uint2(flags[0], flags[1]);
    __syncthreads();  // the words are reused by the next check
    return r;
}

// Per-iteration hard decisions (flow_decoded + K*iteration, ITTC/log_map.cpp:1261-1264): the decision masks of the
// pass that just ran, scattered to rows [row_lo, row_hi) of the codeblock's [n_iter][K] int slab.  A diagnostic output
// (the reference's per-iteration error counting, main.cpp:224-237): plain 4-byte stores, not staged.
__device__ __forceinline__ void emit_iter_bits(const FastArgs &A, const Smem &sm, int NW, int P, int PP, int L, int K, int t, bool natural,
                                               int cbA, bool hasB, int row_lo, int row_hi)
{
    for (int w2 = 0; w2 < (NW + 1) / 2; w2++) {
        const w32 word = sm.dec[w2 * P + t];
#pragma unroll 1
        for (int kk = 0; kk < 16; kk++) {
            const int j = 16 * w2 + (kk & 8) + 7 - (kk & 7);
            if (j >= L) continue;
            int n = t * L + j;
            if (!natural) {
                const int e = sm.tab[j * PP + t];
                const int jj = e / PP, tt = e - jj * PP;
                n = tt * L + jj;
            }
            const int32_t ba = (int32_t)(((word >> kk) & 1u) ^ 1u), bb = (int32_t)(((word >> (16 + kk)) & 1u) ^ 1u);
            for (int r = row_lo; r < row_hi; r++) {
                A.bits_iters[((size_t)cbA * A.n_iter + r) * K + n] = ba;
                if (hasB) A.bits_iters[((size_t)(cbA + 1) * A.n_iter + r) * K + n] = bb;
            }
        }
    }
}

#ifndef TDB_MINB_RT
#define TDB_MINB_RT 2
#endif
#ifndef TDB_MINB_P128_NW4
#define TDB_MINB_P128_NW4 3
#endif
template <int LLR_T, int KP, int KNW, int KG, bool CRC, bool LM = false>
__global__ void __launch_bounds__(KP > 0 ? ((KP + 31) / 32) * 32 : (KP == -1 ? 128 : (KP == -2 ? 192 : 256)), (KP == 128 && KNW == 4 && !LM) ? TDB_MINB_P128_NW4 : ((KP == -1 && !LM) ? TDB_MINB_RT : (KP ? 2 : 1))) fast_s16_kernel(FastArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FastGeom &g = A.g;
    // KP > 0: P, NW, G all compile-time; KP < 0: P at run time (-1: P <= 128, -2: P <= 192), NW and G compile-time; KP == 0: all run-time
    const int P = KP > 0 ? KP : g.P, PP = KP > 0 ? (KP | 1) : g.PP, NW = KP ? KNW : g.NW;
    const int L = 8 * NW, K = P * L;
    const int W = L * PP;            // words per array
    const int Wp = (W + 7) & ~7;     // every region starts 16-byte aligned
    // codeblock pairs this CTA decodes side by side: short blocks, and sub-block counts that would leave
    // most of a CTA's last warp idle (P = 41: three pairs fill 123 of 128 lanes instead of 41 of 64)
    constexpr bool kSingle = KP > 0 || KP == -2;  // kernels that never pack (P > 128 has no room for a second pair)
    const int NP = kSingle ? 1 : A.pairs_per_cta;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n_pairs = (A.n_cb + 1) / 2;
    // thread -> (pair slot q, sub-block t)
    const int q = kSingle ? 0 : tid / P, t = kSingle ? tid : tid - q * P;
    const int pair = blockIdx.x * NP + q;
    const bool active = KP > 0 ? true : (kSingle ? tid < P : (q < NP && pair < n_pairs));
    const bool one_pair = kSingle || NP == 1;  // CTA-uniform
    const Smem sm = pair_smem(smem_raw, g, P, NW, Wp, NP, active ? q : 0);
    unsigned *flags = reinterpret_cast<unsigned *>(sm.edge + 32 * (nthr >> 5));
    const int cbA = 2 * (active ? pair : 0);
    const bool hasB = cbA + 1 < A.n_cb;
    const int cbB = hasB ? cbA + 1 : cbA;
    const size_t row = (size_t)3 * K + 12;
    const float scale = (float)(1 << A.frac_bits);
    const int clip = A.llr_clip;

    // ---- load + quantise + de-multiplex (once per decode); element n = tt*L + j -> word j*PP + tt.
    //      NG groups of loads (NG x 2 codeblocks x 48 bytes) are in flight per thread before the
    //      first is consumed: the phase is pure load latency, so depth is what shortens it.
    //      With several pairs per CTA the groups of all pairs form one index space, so that short blocks
    //      (K/4 groups per pair < CTA size) still keep every thread loading.
    {
        const int np_here = kSingle ? 1 : min(NP, n_pairs - (int)blockIdx.x * NP);
        constexpr int NG = (LLR_T == TDB200_LLR_F64 || LLR_T == kLlrSymQpskF32) ? 2 : 4;
        const w32 clipv = dup2(clip), nclipv = dup2(-clip);
        const int nq = K / 4, total = np_here * nq;
        for (int i0 = tid; i0 < total; i0 += NG * nthr) {
            Raw12<LLR_T> ra[NG], rb[NG];
#pragma unroll
            for (int j = 0; j < NG; j++) {
                const int ii = min(i0 + j * nthr, total - 1);  // the clamp re-reads the last group instead of branching
                const int p = one_pair ? 0 : ii / nq, qq = ii - p * nq;
                const int a_cb = 2 * ((int)blockIdx.x * NP + p), b_cb = (a_cb + 1 < A.n_cb) ? a_cb + 1 : a_cb;
                load12<LLR_T>(A, row, a_cb, qq, ra[j]);
                load12<LLR_T>(A, row, b_cb, qq, rb[j]);
            }
#pragma unroll
            for (int j = 0; j < NG; j++) {
                const int ii = i0 + j * nthr;
                if (ii < total) {
                    const int p = one_pair ? 0 : ii / nq, qq = ii - p * nq;
                    w32 v[12];
                    pack12<LLR_T>(ra[j], rb[j], scale, clipv, nclipv, v, A.kf);
                    put4(pair_smem(smem_raw, g, P, NW, Wp, NP, p), qq, L, PP, v);
                }
            }
        }
    }
    {   // QPP table: 128-bit loads (the table and the shared-memory array are 16-byte aligned)
        const int n16 = (W * 2) / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(A.tab2);
        uint4 *dst = reinterpret_cast<uint4 *>(sm.tab);
        for (int i = tid; i < n16; i += nthr) dst[i] = __ldg(src + i);
        for (int i = n16 * 8 + tid; i < W; i += nthr) sm.tab[i] = __ldg(A.tab2 + i);
    }
    // ---- pull the rows of the codeblock pair that will run on this SM slot next into L2 (bulk
    //      prefetch, a few KB per instruction; rows are 16-byte multiples)
    if (KP && A.prefetch_stride > 0 && LLR_T != kLlrSymQpskF32) {
        const long long nxt = (long long)2 * NP * (blockIdx.x + A.prefetch_stride);
        if (nxt < A.n_cb) {
            const long long nrows = (A.n_cb - nxt < 2 * NP) ? (A.n_cb - nxt) : 2 * NP;
            const size_t esz = LLR_T == TDB200_LLR_F64 ? 8 : ((LLR_T == TDB200_LLR_F32 || LLR_T == kLlrSymBpskF32) ? 4 : (LLR_T == TDB200_LLR_F16 ? 2 : 1));
            // 16-byte granules: with one-byte channel values a row pair starts on an 8-byte boundary
            const size_t b0 = reinterpret_cast<size_t>(A.llr) + (size_t)nxt * row * esz;
            const size_t lo = b0 & ~(size_t)15, hi = (b0 + (size_t)nrows * row * esz) & ~(size_t)15;
            const char *p = reinterpret_cast<const char *>(lo);
            const size_t nbytes = hi - lo;
            const size_t chunk = 4096;
            for (size_t o = (size_t)tid * chunk; o < nbytes; o += (size_t)nthr * chunk) {
                const unsigned sz = (unsigned)(nbytes - o < chunk ? nbytes - o : chunk);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + o), "r"(sz) : "memory");
            }
        }
    }

    PassCfg c;
    c.q2 = A.q2;
    c.lim1 = dup2(A.ext_lim + 1);
    c.limmax = dup2(2 * A.ext_lim - 1);
    c.unbias = dup2((A.q2 == 3 ? -(3 * A.ext_lim / 4) : -A.ext_lim) - 128);
    c.neg1 = A.opaque[0]; c.four = A.opaque[1]; c.k64k = A.opaque[2]; c.k3q = A.opaque[3];
    c.etT = dup2(A.et_threshold); c.et2T = dup2(2 * A.et_threshold); c.etmask = dup2(2 * A.et_threshold - 1);
    c.lk = dup2(1 + A.lm_t4 - 0x4000); c.lku = dup2(1 + A.lm_t4);
    c.lm = 0x3fff3fffu; c.lx2 = 0x20002000u; c.lx1 = 0x1fff1fffu; c.lkc = 0xC000C000u;
    const bool first_fixed = (t == 0), last_fixed = (t == P - 1);

    // ---- boundary vectors.  [s][0..7]: s = SISO
    w32 na[2][8], nb[2][8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            na[s][j] = (first_fixed && j) ? dup2(kFxNeg) : 0u;  // known start state, :943-948
            nb[s][j] = 0u;
        }
        if (last_fixed && active) {
            // termination folded into beta(K): three tail steps back from state 0, :950-954
            w32 bt[8];
#pragma unroll
            for (int j = 0; j < 8; j++) bt[j] = j ? dup2(kFxNeg) : 0u;
            for (int m = 2; m >= 0; m--) {
                const size_t o = (size_t)3 * K + 6 * s + 2 * m;
                const w32 u = pack2(load1<LLR_T>(A, cbA, row, o, scale, clip), load1<LLR_T>(A, cbB, row, o, scale, clip));
                const w32 v = pack2(load1<LLR_T>(A, cbA, row, o + 1, scale, clip), load1<LLR_T>(A, cbB, row, o + 1, scale, clip));
                beta_step_x<LM>(bt, u, v, c);
            }
            norm8(bt);
#pragma unroll
            for (int j = 0; j < 8; j++) { nb[s][j] = bt[j]; sm.tail[8 * s + j] = bt[j]; }
        }
    }
    __syncthreads();

    const bool want_soft = (A.llr2 != nullptr);
    // CRC: the stopping rule by CRC has its own instantiations (the default kernels do not carry the path); one pair per CTA
    bool natural = false;                     // the delivered decisions are SISO-1's (natural order)
    int used = A.n_iter, usedA = 0, usedB = 0;
    if constexpr (LM) {
        // one copy of the pass serves both SISOs: the boundary vectors of the SISO that runs next sit in na[0] / nb[0]
        // and trade places with the other pair after every pass
#pragma unroll 1
        for (int hp = 0; hp < 2 * A.n_iter; hp++) {
            const bool il = (hp & 1) != 0;
            const int it = hp >> 1;
            const bool last = (it == A.n_iter - 1);
            // the CRC stopping rule is a run-time mode of the Log-MAP kernels (the pass has run-time flags anyway): from the
            // second iteration on SISO-1 keeps its natural-order decisions and the pair stops when both codeblocks divide
            const bool crc_pass = (A.early_term == 2) && !il && it >= 1;
            const bool want = (il && (A.early_term == 1 || last || A.bits_iters != nullptr)) || crc_pass;
            w32 weak = 0;
            const w32 chg = siso_pass<-1, -1, KP, KNW, KG, true>(c, g, sm, il ? sm.par2 : sm.par1, na[0], nb[0], t, active, first_fixed, last_fixed,
                                                                 (want_soft && last && il) ? sm.par1 : nullptr, weak,
                                                                 (A.early_term == 2) ? 0xffffffffu : ((usedA ? 0u : 0xffffu) | (usedB ? 0u : 0xffff0000u)), il, want);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const w32 ta = na[0][j], tb2 = nb[0][j];
                na[0][j] = na[1][j]; nb[0][j] = nb[1][j];
                na[1][j] = ta; nb[1][j] = tb2;
            }
            if (il && A.bits_iters && active) emit_iter_bits(A, sm, NW, P, PP, L, K, t, false, cbA, hasB, it, it + 1);
            if (crc_pass) {
                const uint2 rem = crc_of_decisions(A, sm, flags, P, NW, t, active);
                if (!rem.x && !usedA) usedA = it + 1;
                if (!rem.y && !usedB) usedB = it + 1;
                if (usedA && usedB) { used = it + 1; natural = true; break; }
            }
            if (il && A.early_term == 1) {
                int chA, chB;
                if (one_pair) {
                    chA = __syncthreads_or((int)((chg | weak) & 0xffffu));
                    chB = __syncthreads_or((int)((chg | weak) >> 16));
                } else {
                    if (tid < NP) flags[tid] = 0u;
                    __syncthreads();
                    if (active && (chg | weak)) atomicOr(&flags[q], chg | weak);
                    __syncthreads();
                    const unsigned f = active ? flags[q] : 0u;
                    chA = (int)(f & 0xffffu); chB = (int)(f >> 16);
                }
                if (it >= 1) {
                    if (!chA && !usedA) usedA = it + 1;
                    if (!chB && !usedB) usedB = it + 1;
                }
                const bool done = (usedA && usedB) || (!one_pair && !active);
                if (one_pair ? done : (__syncthreads_and((int)done) != 0)) { used = it + 1; break; }
            }
        }
    } else
    for (int it = 0; it < A.n_iter; it++) {
        const bool last = (it == A.n_iter - 1);
        w32 weak = 0;
        bool siso1_done = false;
        if constexpr (CRC) {
            if (it >= 1) {
                // SISO-1 with decisions; stop when the natural-order decisions of both codeblocks divide by the
                // generator -- half an iteration after the SISO-2 pass that made them right
                siso_pass<0, 1, KP, KNW, KG, LM>(c, g, sm, sm.par1, na[0], nb[0], t, active, first_fixed, last_fixed, nullptr, weak);
                const uint2 rem = crc_of_decisions(A, sm, flags, P, NW, t, active);
                if (!rem.x && !usedA) usedA = it + 1;
                if (!rem.y && !usedB) usedB = it + 1;
                if (usedA && usedB) { used = it + 1; natural = true; break; }
                siso1_done = true;
            }
        }
        if (!siso1_done) siso_pass<0, 0, KP, KNW, KG, LM>(c, g, sm, sm.par1, na[0], nb[0], t, active, first_fixed, last_fixed, nullptr, weak);
        if (A.early_term == 1 || last || A.bits_iters != nullptr) {
            // with soft outputs requested, the last SISO-2 pass parks the a-posteriori values in the
            // (by then dead) parity-1 array
            const w32 chg = siso_pass<1, 1, KP, KNW, KG, LM>(c, g, sm, sm.par2, na[1], nb[1], t, active, first_fixed, last_fixed,
                                                               (want_soft && last) ? sm.par1 : nullptr, weak,
                                                               CRC ? 0xffffffffu : ((usedA ? 0u : 0xffffu) | (usedB ? 0u : 0xffff0000u)));
            if (A.bits_iters && active) emit_iter_bits(A, sm, NW, P, PP, L, K, t, false, cbA, hasB, it, it + 1);
            if (A.early_term == 1) {
                // stop: no decision of this iteration differs from the previous one and no
                // a-posteriori value is weaker than the threshold -- per codeblock; a CTA leaves when
                // all its codeblocks have stopped
                int chA, chB;
                if (one_pair) {
                    chA = __syncthreads_or((int)((chg | weak) & 0xffffu));
                    chB = __syncthreads_or((int)((chg | weak) >> 16));
                } else {
                    if (tid < NP) flags[tid] = 0u;
                    __syncthreads();
                    if (active && (chg | weak)) atomicOr(&flags[q], chg | weak);
                    __syncthreads();
                    const unsigned f = active ? flags[q] : 0u;
                    chA = (int)(f & 0xffffu); chB = (int)(f >> 16);
                }
                if (it >= 1) {
                    if (!chA && !usedA) usedA = it + 1;
                    if (!chB && !usedB) usedB = it + 1;
                }
                const bool done = (usedA && usedB) || (!one_pair && !active);  // with one pair per CTA the flags are CTA-uniform
                if (one_pair ? done : (__syncthreads_and((int)done) != 0)) { used = it + 1; break; }
            }
        } else {
            siso_pass<1, 0, KP, KNW, KG, LM>(c, g, sm, sm.par2, na[1], nb[1], t, active, first_fixed, last_fixed, nullptr, weak);
        }
    }
    if (!usedA) usedA = used;
    if (!usedB) usedB = used;
    if (A.bits_iters && active && (used < A.n_iter || natural))  // rows past an early stop repeat the last one
        emit_iter_bits(A, sm, NW, P, PP, L, K, t, natural, cbA, hasB, natural ? used - 1 : used, A.n_iter);

    // ---- hard decisions, natural order: decision() :862-879 is the sign bit kept in sm.dec,
    //      random_deinterlvr_int :1264 is the scatter of those bits to byte n = pi(i) of a staging
    //      array (the dead parity-2 region), which then leaves with coalesced 32-bit stores
    if (A.bits) {
        if (active) {
            uint8_t *byA = reinterpret_cast<uint8_t *>(sm.par2), *byB = byA + K;
            for (int w2 = 0; w2 < (NW + 1) / 2; w2++) {
                const w32 word = sm.dec[w2 * P + t];
#pragma unroll
                for (int kk = 0; kk < 16; kk++) {
                    const int j = 16 * w2 + (kk & 8) + 7 - (kk & 7);  // step k of a window sits in bit 7-k of its byte
                    if (j < L) {
                        int n = t * L + j;  // SISO-1 decisions are in natural order already
                        if (!natural) {
                            const int e = sm.tab[j * PP + t];
                            const int jj = e / PP, tt = e - jj * PP;
                            n = tt * L + jj;
                        }
                        byA[n] = (uint8_t)(((word >> kk) & 1u) ^ 1u);
                        byB[n] = (uint8_t)(((word >> (16 + kk)) & 1u) ^ 1u);
                    }
                }
            }
        }
        __syncthreads();
        {   // the pairs of a CTA are consecutive rows of the output: one index space over all of them
            const int np_here = kSingle ? 1 : min(NP, n_pairs - (int)blockIdx.x * NP);
            const int nw = K / 4, total = np_here * nw;
            for (int ii = tid; ii < total; ii += nthr) {
                const int p = one_pair ? 0 : ii / nw, i = ii - p * nw;
                const int pr = blockIdx.x * NP + p;
                const uint32_t *wa = reinterpret_cast<const uint32_t *>(pair_smem(smem_raw, g, P, NW, Wp, NP, p).par2), *wb = wa + nw;
                uint32_t *oa = reinterpret_cast<uint32_t *>(A.bits + (size_t)(2 * pr) * K), *ob = oa + nw;
                oa[i] = wa[i];
                if (2 * pr + 1 < A.n_cb) ob[i] = wb[i];
            }
        }
    }
    if (A.iters_used && active && t == 0) {
        A.iters_used[cbA] = usedA;
        if (hasB) A.iters_used[cbB] = usedB;
    }
    if (want_soft || A.ext2) {
        const float inv = 1.0f / scale;
        const int T = K + kTail;
        __syncthreads();
        for (int p = 0; p < NP; p++) {
            const int pr = blockIdx.x * NP + p;
            if (pr >= n_pairs) break;
            const Smem smp = pair_smem(smem_raw, g, P, NW, Wp, NP, p);
            const int a_cb = 2 * pr;
            const bool pB = a_cb + 1 < A.n_cb;
            for (int i = tid; i < T; i += nthr) {
                float la = 0.f, lb = 0.f, ea = 0.f, eb = 0.f;
                if (i < K) {
                    const int tt = i / L, j = i - tt * L;
                    const int e = smp.tab[j * PP + tt];
                    const w32 lam = smp.par1[e];
                    const w32 ysb = (w32)smp.sysA[e] | ((w32)smp.sysB[e] << 16);
                    const w32 ex = vadd(vadd(smp.X[e], vneg(ysb)), 0x00800080u);
                    la = (float)(int16_t)(lam & 0xffff) * inv; lb = (float)(int16_t)(lam >> 16) * inv;
                    ea = (float)(int16_t)(ex & 0xffff) * inv; eb = (float)(int16_t)(ex >> 16) * inv;
                }
                if (A.llr2) { A.llr2[(size_t)a_cb * T + i] = la; if (pB) A.llr2[(size_t)(a_cb + 1) * T + i] = lb; }
                if (A.ext2) { A.ext2[(size_t)a_cb * T + i] = ea; if (pB) A.ext2[(size_t)(a_cb + 1) * T + i] = eb; }
            }
        }
    }
}

typedef void (*kernel_fn)(FastArgs);

template <int LLR_T, int KP, bool CRC>
kernel_fn pick_nw(int NW)
{
    if (NW == 6) return fast_s16_kernel<LLR_T, KP, 6, 16, CRC>;
    if (NW == 5) return fast_s16_kernel<LLR_T, KP, 5, 16, CRC>;
    return fast_s16_kernel<LLR_T, KP, 4, 16, CRC>;
}

template <int LLR_T, bool CRC>
kernel_fn pick_rt(int NW)
{
    switch (NW) {
        case 4: return fast_s16_kernel<LLR_T, -1, 4, 16, CRC>;
        case 5: return fast_s16_kernel<LLR_T, -1, 5, 16, CRC>;
        case 6: return fast_s16_kernel<LLR_T, -1, 6, 16, CRC>;
        case 7: return fast_s16_kernel<LLR_T, -1, 7, 16, CRC>;
        default: return fast_s16_kernel<LLR_T, -1, 8, 16, CRC>;
    }
}

template <int LLR_T, bool CRC>
kernel_fn pick_kernel_t(const FastGeom &g)
{
    if (fast_spec_pn(g)) return g.P == 128 ? pick_nw<LLR_T, 128, CRC>(g.NW) : (g.P == 64 ? pick_nw<LLR_T, 64, CRC>(g.NW) : pick_nw<LLR_T, 32, CRC>(g.NW));
    if (fast_spec128g8(g)) return fast_s16_kernel<LLR_T, 128, 6, 8, CRC>;
    if (fast_spec192(g)) return fast_s16_kernel<LLR_T, 192, 4, 16, CRC>;
    if (fast_spec_rt(g)) return pick_rt<LLR_T, CRC>(g.NW);
    if (fast_spec_rt192(g)) return g.NW == 4 ? fast_s16_kernel<LLR_T, -2, 4, 16, CRC> : fast_s16_kernel<LLR_T, -2, 5, 16, CRC>;
    return fast_s16_kernel<LLR_T, 0, 0, 0, CRC>;
}

// TDB200_ALGO_LOGMAP_S16: compile-time geometry for 128 sub-blocks of 32 / 40 / 48 steps with guard 16, 24 or 32
// (K = 4096, 5120, 6144); every other plan runs the instantiation with run-time geometry
template <int LLR_T, int G>
kernel_fn pick_lm_nw(int NW)
{
    return NW == 6 ? fast_s16_kernel<LLR_T, 128, 6, G, false, true> : (NW == 5 ? fast_s16_kernel<LLR_T, 128, 5, G, false, true> : fast_s16_kernel<LLR_T, 128, 4, G, false, true>);
}
template <int LLR_T>
kernel_fn pick_lm_rt(int NW)
{
    switch (NW) {
        case 4: return fast_s16_kernel<LLR_T, -1, 4, 24, false, true>;
        case 5: return fast_s16_kernel<LLR_T, -1, 5, 24, false, true>;
        case 6: return fast_s16_kernel<LLR_T, -1, 6, 24, false, true>;
        case 7: return fast_s16_kernel<LLR_T, -1, 7, 24, false, true>;
        default: return fast_s16_kernel<LLR_T, -1, 8, 24, false, true>;
    }
}
template <int LLR_T>
kernel_fn pick_kernel_lm_t(const FastGeom &g)
{
    if (fast_spec_lm(g)) return g.G == 32 ? pick_lm_nw<LLR_T, 32>(g.NW) : (g.G == 24 ? pick_lm_nw<LLR_T, 24>(g.NW) : pick_lm_nw<LLR_T, 16>(g.NW));
    if (fast_spec_lm_rt(g)) return pick_lm_rt<LLR_T>(g.NW);
    return fast_s16_kernel<LLR_T, 0, 0, 0, false, true>;
}

}  // namespace
}  // namespace tdb200
