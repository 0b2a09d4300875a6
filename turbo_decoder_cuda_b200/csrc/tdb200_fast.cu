// tdb200_fast.cu -- TDB200_ALGO_MAXLOG_S16: the throughput decoder.
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
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

typedef uint32_t w32;  // two int16 lanes: codeblock A in bits 0-15, codeblock B in bits 16-31

__device__ __forceinline__ w32 vadd(w32 a, w32 b) { return __vadd2(a, b); }                    // VIADD.16x2  (fma pipe)
__device__ __forceinline__ w32 vaddmax(w32 a, w32 b, w32 c) { return __viaddmax_s16x2(a, b, c); }  // VIADDMNMX.S16x2: max(a+b,c)
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

// alpha(i+1) from alpha(i); u = U_i, v = V_i   (restates the max-log form of :975-1001)
__device__ __forceinline__ void alpha_step(w32 (&a)[8], w32 u, w32 v)
{
    const w32 w = vadd(u, v);
    const w32 t5 = vadd(a[2], v), t1 = vadd(a[3], v), t2 = vadd(a[4], v), t6 = vadd(a[5], v);
    const w32 o0 = vaddmax(a[1], w, a[0]), o4 = vaddmax(a[0], w, a[1]);
    const w32 o5 = vaddmax(a[3], u, t5), o1 = vaddmax(a[2], u, t1);
    const w32 o2 = vaddmax(a[5], u, t2), o6 = vaddmax(a[4], u, t6);
    const w32 o7 = vaddmax(a[7], w, a[6]), o3 = vaddmax(a[6], w, a[7]);
    a[0] = o0; a[1] = o1; a[2] = o2; a[3] = o3; a[4] = o4; a[5] = o5; a[6] = o6; a[7] = o7;
}

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

// e = max_{input 1}(alpha + c*V + beta') - max_{input 0}(alpha + c*V + beta')   (:1024-1039 as
// max-log; the +U common to all input-1 branches is left out, so e IS the extrinsic :1234-1238)
__device__ __forceinline__ w32 extrinsic(const w32 (&a)[8], const w32 (&b)[8], w32 v)
{
    w32 m0a = vadd(a[0], b[0]); m0a = vaddmax(a[1], b[4], m0a); m0a = vaddmax(a[6], b[7], m0a); m0a = vaddmax(a[7], b[3], m0a);
    w32 m0b = vadd(a[2], b[5]); m0b = vaddmax(a[3], b[1], m0b); m0b = vaddmax(a[4], b[2], m0b); m0b = vaddmax(a[5], b[6], m0b);
    w32 m1a = vadd(a[0], b[4]); m1a = vaddmax(a[1], b[0], m1a); m1a = vaddmax(a[6], b[3], m1a); m1a = vaddmax(a[7], b[7], m1a);
    w32 m1b = vadd(a[2], b[1]); m1b = vaddmax(a[3], b[5], m1b); m1b = vaddmax(a[4], b[6], m1b); m1b = vaddmax(a[5], b[2], m1b);
    const w32 m0 = vaddmax(m0b, v, m0a);
    const w32 m1 = vaddmax(m1a, v, m1b);
    return vadd(m1, vneg(m0));
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

// 12 consecutive input values (4 systematic/parity1/parity2 triplets) starting at element 12*q
__device__ __forceinline__ void load12(const void *base, int type, size_t row_elems, int cb, int q, float scale, int clip, int (&out)[12])
{
    if (type == TDB200_LLR_F32) {
        const float4 *p = reinterpret_cast<const float4 *>(static_cast<const float *>(base) + (size_t)cb * row_elems) + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float4 f = __ldg(p + k);
            out[4 * k] = quant(f.x, scale, clip); out[4 * k + 1] = quant(f.y, scale, clip);
            out[4 * k + 2] = quant(f.z, scale, clip); out[4 * k + 3] = quant(f.w, scale, clip);
        }
    } else if (type == TDB200_LLR_F64) {
        const double2 *p = reinterpret_cast<const double2 *>(static_cast<const double *>(base) + (size_t)cb * row_elems) + 6 * q;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double2 f = __ldg(p + k);
            out[2 * k] = quant((float)f.x, scale, clip); out[2 * k + 1] = quant((float)f.y, scale, clip);
        }
    } else {
        const int *p = reinterpret_cast<const int *>(static_cast<const int8_t *>(base) + (size_t)cb * row_elems) + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int wv = __ldg(p + k);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int v = (int)(int8_t)((wv >> (8 * m)) & 0xff);
                out[4 * k + m] = max(min(v, clip), -clip);
            }
        }
    }
}

__device__ __forceinline__ int load1(const void *base, int type, size_t idx, float scale, int clip)
{
    if (type == TDB200_LLR_F32) return quant(__ldg(static_cast<const float *>(base) + idx), scale, clip);
    if (type == TDB200_LLR_F64) return quant((float)__ldg(static_cast<const double *>(base) + idx), scale, clip);
    const int v = (int)__ldg(static_cast<const int8_t *>(base) + idx);
    return max(min(v, clip), -clip);
}

struct Smem {
    w32 *X, *par1, *par2;
    uint16_t *sys8, *tab;
    w32 *ckpt, *edge;
};

__device__ __forceinline__ Smem carve(unsigned char *base, int K, int P, int n_ckpt)
{
    Smem s;
    s.X = reinterpret_cast<w32 *>(base);
    s.par1 = s.X + K;
    s.par2 = s.par1 + K;
    s.sys8 = reinterpret_cast<uint16_t *>(s.par2 + K);
    s.tab = s.sys8 + K;
    s.ckpt = reinterpret_cast<w32 *>(s.tab + K);
    s.edge = s.ckpt + (size_t)n_ckpt * 7 * P;
    return s;
}

// sign-extend the two int8 of a 16-bit word into an s16x2
__device__ __forceinline__ w32 sext8x2(unsigned v)
{
    w32 r;  // prmt default mode: selector nibble bit 3 replicates the sign of the selected byte
    asm("prmt.b32 %0, %1, 0, 0x9180;" : "=r"(r) : "r"(v));
    return r;
}

struct PassCfg {
    int t, P, L, NW, G, n_ckpt;
    int q2;
    w32 lim;       // dup2(ext_lim)
    w32 limmax;    // dup2(2*ext_lim - 1)
    w32 unbias;    // dup2(-(3*ext_lim/4)) or dup2(-ext_lim)
};

// One SISO pass of one sub-block.  IL = false: SISO-1 (natural order), true: SISO-2 (through tab).
// na/nb: boundary vectors (alpha G steps before the sub-block, beta G steps after it); on return
// they hold the vectors for the next iteration of this SISO.
template <bool IL>
__device__ __forceinline__ void siso_pass(const PassCfg &c, const Smem &sm, const w32 *par, w32 (&na)[8], w32 (&nb)[8],
                                          const bool first_fixed, const bool last_fixed, w32 *stage)
{
    const int t = c.t, P = c.P, NW = c.NW, G = c.G;
    const bool active = t < P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    w32 a[8], b[8], a0[8], sa[8], sb[8];
#pragma unroll
    for (int s = 0; s < 8; s++) { a[s] = na[s]; b[s] = nb[s]; sa[s] = 0; sb[s] = 0; }

    if (active) {
        // ---- alpha warm-up over the last G steps of sub-block t-1
        if (!first_fixed)
            for (int g0 = 0; g0 < G; g0 += 8) {
                norm8(a);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int idx = (c.L - G + g0 + k) * P + (t - 1);
                    const int e = IL ? sm.tab[idx] : idx;
                    alpha_step(a, sm.X[e], par[idx]);
                }
            }
        // ---- beta warm-up over the first G steps of sub-block t+1
        if (!last_fixed)
            for (int g0 = G - 8; g0 >= 0; g0 -= 8) {
                norm8(b);
#pragma unroll
                for (int k = 7; k >= 0; k--) {
                    const int idx = (g0 + k) * P + (t + 1);
                    const int e = IL ? sm.tab[idx] : idx;
                    beta_step(b, sm.X[e], par[idx]);
                }
            }
        if (G == c.L) {
#pragma unroll
            for (int s = 0; s < 8; s++) sb[s] = b[s];
        }
#pragma unroll
        for (int s = 0; s < 8; s++) a0[s] = a[s];
        // ---- forward sweep over windows 0..NW-2, leaving a checkpoint at the start of windows 1..NW-2
        for (int w = 0; w < NW - 1; w++) {
            norm8(a);
            if (w > 0) {
#pragma unroll
                for (int s = 1; s < 8; s++) sm.ckpt[((w - 1) * 7 + (s - 1)) * P + t] = a[s];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int idx = (8 * w + k) * P + t;
                const int e = IL ? sm.tab[idx] : idx;
                alpha_step(a, sm.X[e], par[idx]);
            }
        }
    }
    __syncthreads();  // every warm-up read of X precedes every in-place update below
    if (active) {
        const int w_sa = (c.L - G) >> 3, w_sb = G >> 3;
        for (int w = NW - 1; w >= 0; w--) {
            if (w < NW - 1) {
                if (w == 0) {
#pragma unroll
                    for (int s = 0; s < 8; s++) a[s] = a0[s];
                } else {
                    a[0] = 0;
#pragma unroll
                    for (int s = 1; s < 8; s++) a[s] = sm.ckpt[((w - 1) * 7 + (s - 1)) * P + t];
                }
            }
            // ---- re-create the 8 alpha vectors of this window in registers
            w32 aw[8][8], u[8], v[8];
            int e[8];
            norm8(a);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int idx = (8 * w + k) * P + t;
                e[k] = IL ? sm.tab[idx] : idx;
                u[k] = sm.X[e[k]];
                v[k] = par[idx];
#pragma unroll
                for (int s = 0; s < 8; s++) aw[k][s] = a[s];
                alpha_step(a, u[k], v[k]);
            }
            if (w == w_sa) {
#pragma unroll
                for (int s = 0; s < 8; s++) sa[s] = aw[0][s];
            }
            if (G == 0 && w == NW - 1) {
#pragma unroll
                for (int s = 0; s < 8; s++) sa[s] = a[s];
            }
            // ---- beta, extrinsic, in-place update of X
            norm8(b);
#pragma unroll
            for (int k = 7; k >= 0; k--) {
                const w32 ex = extrinsic(aw[k], b, v[k]);
                if (stage) stage[e[k]] = vadd(u[k], ex);  // a-posteriori, :1038 (+ the dropped U)
                // clamp to [-lim, lim-1], bias to [0, 2lim-1]
                const w32 y = __viaddmin_s16x2_relu(ex, c.lim, c.limmax);
                w32 es;
                if (c.q2 == 3) es = ((y * 3u) >> 2) & 0x3fff3fffu;  // floor(3(ec+lim)/4), no cross-lane carry
                else es = y;
                const w32 ys = sext8x2(sm.sys8[e[k]]);
                sm.X[e[k]] = vadd(vadd(ys, es), c.unbias);
                beta_step(b, u[k], v[k]);
            }
            if (w == w_sb) {
#pragma unroll
                for (int s = 0; s < 8; s++) sb[s] = b[s];
            }
        }
        norm8(sa);
        norm8(sb);
    }
    // ---- hand the boundary vectors to the neighbours (they use them in the next iteration)
    w32 up[8], dn[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        up[s] = __shfl_up_sync(0xffffffffu, sa[s], 1);
        dn[s] = __shfl_down_sync(0xffffffffu, sb[s], 1);
        if (lane == 31) sm.edge[s * nwarps + warp] = sa[s];
        if (lane == 0) sm.edge[(8 + s) * nwarps + warp] = sb[s];
    }
    __syncthreads();  // also orders this pass's X updates before the next pass's reads
#pragma unroll
    for (int s = 0; s < 8; s++) {
        if (lane == 0 && warp > 0) up[s] = sm.edge[s * nwarps + warp - 1];
        if (lane == 31 && warp + 1 < nwarps) dn[s] = sm.edge[(8 + s) * nwarps + warp + 1];
        if (!first_fixed) na[s] = up[s];
        if (!last_fixed) nb[s] = dn[s];
    }
}

__global__ void __launch_bounds__(256) fast_s16_kernel(FastArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FastGeom &g = A.g;
    const int K = g.K, L = g.L, P = g.P;
    const Smem sm = carve(smem_raw, K, P, g.n_ckpt);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int cbA = 2 * blockIdx.x;
    const bool hasB = cbA + 1 < A.n_cb;
    const int cbB = hasB ? cbA + 1 : cbA;
    const size_t row = (size_t)3 * K + 12;
    const float scale = (float)(1 << A.frac_bits);
    const int clip = A.llr_clip;

    // ---- load + quantise + de-multiplex (once per decode)
    for (int q = tid; q < K / 4; q += nthr) {
        int va[12], vb[12];
        load12(A.llr, A.llr_type, row, cbA, q, scale, clip, va);
        load12(A.llr, A.llr_type, row, cbB, q, scale, clip, vb);
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int n = 4 * q + m;
            const int tt = n / L, j = n - tt * L;
            const int ad = j * P + tt;
            sm.X[ad] = pack2(va[3 * m], vb[3 * m]);
            sm.sys8[ad] = (uint16_t)((va[3 * m] & 0xff) | ((vb[3 * m] & 0xff) << 8));
            sm.par1[ad] = pack2(va[3 * m + 1], vb[3 * m + 1]);
            sm.par2[ad] = pack2(va[3 * m + 2], vb[3 * m + 2]);
        }
    }
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.tab2);
        uint32_t *dst = reinterpret_cast<uint32_t *>(sm.tab);
        for (int i = tid; i < K / 2; i += nthr) dst[i] = __ldg(src + i);
    }

    PassCfg c;
    c.t = tid; c.P = P; c.L = L; c.NW = g.NW; c.G = g.G; c.n_ckpt = g.n_ckpt; c.q2 = A.q2;
    c.lim = dup2(A.ext_lim);
    c.limmax = dup2(2 * A.ext_lim - 1);
    c.unbias = dup2(A.q2 == 3 ? -(3 * A.ext_lim / 4) : -A.ext_lim);
    const bool first_fixed = (tid == 0), last_fixed = (tid == P - 1);

    // ---- boundary vectors.  [s][0..7]: s = SISO
    w32 na[2][8], nb[2][8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            na[s][j] = (first_fixed && j) ? dup2(kFxNeg) : 0u;  // known start state, :943-948
            nb[s][j] = 0u;
        }
        if (last_fixed) {
            // termination folded into beta(K): three tail steps back from state 0, :950-954
            w32 bt[8];
#pragma unroll
            for (int j = 0; j < 8; j++) bt[j] = j ? dup2(kFxNeg) : 0u;
            for (int m = 2; m >= 0; m--) {
                const size_t o = (size_t)3 * K + 6 * s + 2 * m;
                const w32 u = pack2(load1(A.llr, A.llr_type, cbA * row + o, scale, clip), load1(A.llr, A.llr_type, cbB * row + o, scale, clip));
                const w32 v = pack2(load1(A.llr, A.llr_type, cbA * row + o + 1, scale, clip), load1(A.llr, A.llr_type, cbB * row + o + 1, scale, clip));
                beta_step(bt, u, v);
            }
            norm8(bt);
#pragma unroll
            for (int j = 0; j < 8; j++) nb[s][j] = bt[j];
        }
    }
    __syncthreads();

    const bool want_soft = (A.llr2 != nullptr);
    for (int it = 0; it < A.n_iter; it++) {
        const bool last = (it == A.n_iter - 1);
        siso_pass<false>(c, sm, sm.par1, na[0], nb[0], first_fixed, last_fixed, nullptr);
        // the last SISO-2 pass parks the a-posteriori values in the (now dead) parity-1 array
        siso_pass<true>(c, sm, sm.par2, na[1], nb[1], first_fixed, last_fixed, last ? sm.par1 : nullptr);
    }

    // ---- hard decisions, natural order (decision() :862-879 + random_deinterlvr_int :1264 are
    //      the sign bit of the parked value at the bit's own word)
    if (A.bits) {
        for (int q = tid; q < K / 4; q += nthr) {
            uint32_t ba = 0, bb = 0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int n = 4 * q + m;
                const int tt = n / L, j = n - tt * L;
                const w32 lam = sm.par1[j * P + tt];
                ba |= ((lam & 0x8000u) ? 0u : 1u) << (8 * m);
                bb |= ((lam & 0x80000000u) ? 0u : 1u) << (8 * m);
            }
            reinterpret_cast<uint32_t *>(A.bits + (size_t)cbA * K)[q] = ba;
            if (hasB) reinterpret_cast<uint32_t *>(A.bits + (size_t)cbB * K)[q] = bb;
        }
    }
    if (A.iters_used && tid == 0) {
        A.iters_used[cbA] = A.n_iter;
        if (hasB) A.iters_used[cbB] = A.n_iter;
    }
    if (want_soft || A.ext2) {
        const float inv = 1.0f / scale;
        const int T = K + kTail;
        for (int i = tid; i < T; i += nthr) {
            float la = 0.f, lb = 0.f, ea = 0.f, eb = 0.f;
            if (i < K) {
                const int tt = i / L, j = i - tt * L;
                const int e = sm.tab[j * P + tt];
                const w32 lam = sm.par1[e];
                const w32 ex = vadd(sm.X[e], vneg(sext8x2(sm.sys8[e])));
                la = (float)(int16_t)(lam & 0xffff) * inv; lb = (float)(int16_t)(lam >> 16) * inv;
                ea = (float)(int16_t)(ex & 0xffff) * inv; eb = (float)(int16_t)(ex >> 16) * inv;
            }
            if (A.llr2) { A.llr2[(size_t)cbA * T + i] = la; if (hasB) A.llr2[(size_t)cbB * T + i] = lb; }
            if (A.ext2) { A.ext2[(size_t)cbA * T + i] = ea; if (hasB) A.ext2[(size_t)cbB * T + i] = eb; }
        }
    }
}

}  // namespace

int fast_s16_smem_bytes(const FastGeom &g)
{
    const int nwarps = g.threads / 32;
    return 3 * 4 * g.K + 2 * 2 * g.K + 4 * g.n_ckpt * 7 * g.P + 4 * 16 * nwarps;
}

cudaError_t fast_s16_configure(const FastGeom &g)
{
    return cudaFuncSetAttribute(fast_s16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes);
}

cudaError_t launch_fast_s16(const FastArgs &a, cudaStream_t st, int *n_launches)
{
    const int pairs = (a.n_cb + 1) / 2;
    fast_s16_kernel<<<pairs, a.g.threads, a.g.smem_bytes, st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

}  // namespace tdb200
