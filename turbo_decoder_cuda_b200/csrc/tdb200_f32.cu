// tdb200_f32.cu -- TDB200_ALGO_LOGMAP_F32 / LINLOGMAP_F32 / MAXLOG_F32: the sub-block-parallel decoder
// in fp32, with the exact Jacobian correction  max*(x,y) = max(x,y) + ln(1 + e^-|x-y|)  (Log-MAP), its
// linear approximation, or without it (max-log-MAP).
//
// What it computes: the iterative PCCC decode of TurboDecoding() (ITTC/log_map.cpp:1146-1280) with
// the component decoder Log_MAP_decoder() (:898-1047), where the reference's max* (E_algorithm,
// :779-801: max + a 16-step LUT of the same correction) is evaluated exactly with ex2/lg2, and the
// unsegmented recursions are cut into sub-blocks with boundary-state initialisation -- the windowed
// Log-MAP of BASELINE configs[2].  Specification (same schedule, plain C floats):
// oracle/turbo_oracle_f32.c; the max-log variant is bit-exact against it, the Log-MAP variant agrees
// to the accuracy of the hardware ex2/lg2 approximations.
//
// Layout and schedule are those of the packed-int16 kernel (tdb200_fast_kernel.cuh), one codeblock per CTA:
//   * thread t owns trellis steps [tL,(t+1)L), 8 state metrics in registers;
//   * boundary vectors = the neighbour's vector G steps before/after the boundary from the previous
//     iteration, re-run over the G guard steps (warm-up);
//   * alpha checkpoints every 8 steps in shared memory, alpha re-created per window in registers;
//   * gamma(b,c) = b*U + c*V with U = Ls + La kept in place in a step-major array X, V = Lp;
//     SISO-2 reaches X through the QPP table; each SISO overwrites X with Ls + q*Le (q = 1 for
//     Log-MAP as in the reference :1234-1238, 3/4 for max-log);
//   * channel values are stored as fp16 in shared memory (two CTAs per SM at K = 6144); X and all
//     metrics are fp32.
// Compile with -fmad=false: every product here is followed by an add, and the specification does
// not contract them.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

constexpr float kNegF = -1.0e9f;  // metric of an impossible state

// LM: 0 = max, 1 = max* with the exact correction, 2 = max* with the linear correction
//     max(0, 0.24904 * (2.5068 - d))  (least-squares fit of ln(1+e^-d); no special-function unit needed)
template <int LOGMAP>
__device__ __forceinline__ float mx(float x, float y)
{
    const float m = fmaxf(x, y);
    if (LOGMAP == 0) return m;
    if (LOGMAP == 2) return m + fmaxf(__fmaf_rn(-0.24904f, fabsf(x - y), 0.62429345f), 0.f);  // one FFMA, as in the model
    // ln(1 + e^-d) = ln2 * lg2(1 + ex2(-d * log2 e)),  d = |x - y|
    float e, l;
    const float d = fabsf(x - y) * -1.4426950408889634f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(d));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return m + l * 0.6931471805599453f;
}

__device__ __forceinline__ void norm8(float (&m)[8])
{
    const float z = m[0];
    m[0] = 0.f;
#pragma unroll
    for (int s = 1; s < 8; s++) m[s] -= z;
}

// alpha(i+1) from alpha(i)   (:975-1001)
template <int LM>
__device__ __forceinline__ void alpha_step_to(const float (&a)[8], float u, float v, float (&o)[8])
{
    const float w = u + v;
    const float o0 = mx<LM>(a[0], a[1] + w), o4 = mx<LM>(a[0] + w, a[1]);
    const float o5 = mx<LM>(a[2] + v, a[3] + u), o1 = mx<LM>(a[3] + v, a[2] + u);
    const float o2 = mx<LM>(a[4] + v, a[5] + u), o6 = mx<LM>(a[5] + v, a[4] + u);
    const float o7 = mx<LM>(a[6], a[7] + w), o3 = mx<LM>(a[7], a[6] + w);
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5; o[6] = o6; o[7] = o7;
}
template <int LM>
__device__ __forceinline__ void alpha_step(float (&a)[8], float u, float v) { alpha_step_to<LM>(a, u, v, a); }

// beta(i) from beta(i+1)   (:1004-1021)
template <int LM>
__device__ __forceinline__ void beta_step(float (&b)[8], float u, float v)
{
    const float w = u + v;
    const float o0 = mx<LM>(b[0], b[4] + w), o1 = mx<LM>(b[4], b[0] + w);
    const float o2 = mx<LM>(b[5] + v, b[1] + u), o3 = mx<LM>(b[1] + v, b[5] + u);
    const float o4 = mx<LM>(b[2] + v, b[6] + u), o5 = mx<LM>(b[6] + v, b[2] + u);
    const float o6 = mx<LM>(b[7], b[3] + w), o7 = mx<LM>(b[3], b[7] + w);
    b[0] = o0; b[1] = o1; b[2] = o2; b[3] = o3; b[4] = o4; b[5] = o5; b[6] = o6; b[7] = o7;
}

// extrinsic = max*_{input 1}(alpha + c*V + beta') - max*_{input 0}(alpha + c*V + beta')   (:1024-1039
// without the +U common to the input-1 branches, so this IS Le of :1234-1238)
template <int LM>
__device__ __forceinline__ float extrinsic(const float (&a)[8], const float (&b)[8], float v)
{
    const float m0a = mx<LM>(mx<LM>(a[0] + b[0], a[1] + b[4]), mx<LM>(a[6] + b[7], a[7] + b[3]));
    const float m0b = mx<LM>(mx<LM>(a[2] + b[5], a[3] + b[1]), mx<LM>(a[4] + b[2], a[5] + b[6]));
    const float m1a = mx<LM>(mx<LM>(a[0] + b[4], a[1] + b[0]), mx<LM>(a[6] + b[3], a[7] + b[7]));
    const float m1b = mx<LM>(mx<LM>(a[2] + b[1], a[3] + b[5]), mx<LM>(a[4] + b[6], a[5] + b[2]));
    return mx<LM>(m1a + v, m1b) - mx<LM>(m0a, m0b + v);
}

struct Smem {
    float *X;
    __half *sys, *par1, *par2;
    uint16_t *tab;
    float *ckpt;
    uint8_t *dec;  // [NW][P]: decision bits of a window, step k in bit k
    float *edge;
};

struct Pass {
    int P, PP, NW, G, L;
    float q, clamp, etT;
};

template <int LLR_T>
__device__ __forceinline__ float load_llr(const void *base, size_t idx)
{
    float x;
    if (LLR_T == TDB200_LLR_F32) x = __ldg(static_cast<const float *>(base) + idx);
    else if (LLR_T == TDB200_LLR_F64) x = (float)__ldg(static_cast<const double *>(base) + idx);
    else if (LLR_T == TDB200_LLR_F16) x = __half2float(__ldg(static_cast<const __half *>(base) + idx));
    else x = (float)__ldg(static_cast<const int8_t *>(base) + idx) * 0.125f;  // S8: 3 fractional bits
    // NaN -> erasure; the channel values live in shared memory as binary16, so clamp to its finite range first
    // (an LLR beyond +-65504 would become +-inf and inf - inf in the recursions NaN)
    return (x == x) ? fminf(fmaxf(x, -65504.0f), 65504.0f) : 0.f;
}

// One SISO pass of one sub-block; see siso_pass in tdb200_fast_kernel.cuh for the schedule.
template <int LM, bool IL>
__device__ __forceinline__ unsigned siso_pass(const Pass &c, const Smem &sm, const __half *par, float (&na)[8], float (&nb)[8],
                                              const bool first_fixed, const bool last_fixed, const bool want, float *g_llr, float *g_ext,
                                              bool &weak)
{
    const int P = c.P, PP = c.PP, NW = c.NW, G = c.G, L = c.L;
    const int t = threadIdx.x;
    const bool active = t < P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int w_sa = (L - G) >> 3, w_sb = G >> 3;
    float a[8], b[8], a0[8], sa[8], sb[8];
    unsigned changed = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) { a[s] = na[s]; b[s] = nb[s]; sa[s] = 0.f; sb[s] = 0.f; }
    auto elem = [&](int idx) -> int { return IL ? (int)sm.tab[idx] : idx; };

    if (active) {
        const int ta = first_fixed ? t : t - 1, tb = last_fixed ? t : t + 1;
#pragma unroll 1
        for (int g0 = 0; g0 < G; g0 += 8) {
            const int base_a = (L - G + g0) * PP + ta, base_b = (G - 8 - g0) * PP + tb;
            norm8(a);
            norm8(b);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int ia = base_a + k * PP, ib = base_b + (7 - k) * PP;
                alpha_step<LM>(a, sm.X[elem(ia)], __half2float(par[ia]));
                beta_step<LM>(b, sm.X[elem(ib)], __half2float(par[ib]));
            }
        }
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if (first_fixed) a[s] = na[s];
            if (last_fixed) b[s] = nb[s];
        }
        if (G == L) {
#pragma unroll
            for (int s = 0; s < 8; s++) sb[s] = b[s];
        }
        norm8(a);
#pragma unroll
        for (int s = 0; s < 8; s++) a0[s] = a[s];
#pragma unroll 1
        for (int w = 0; w < NW - 1; w++) {
            if (w > 0) {
                norm8(a);
#pragma unroll
                for (int s = 1; s < 8; s++) sm.ckpt[((w - 1) * 7 + (s - 1)) * P + t] = a[s];
            }
            const int base = 8 * w * PP + t;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int idx = base + k * PP;
                alpha_step<LM>(a, sm.X[elem(idx)], __half2float(par[idx]));
            }
        }
        if (NW > 1) norm8(a);
    }
    __syncthreads();  // every warm-up read of X precedes every in-place update below
    if (active) {
        if (G > 0) {
            if (w_sa == NW - 1) {
#pragma unroll
                for (int s = 0; s < 8; s++) sa[s] = a[s];
            } else if (w_sa == 0) {
#pragma unroll
                for (int s = 0; s < 8; s++) sa[s] = a0[s];
            }
        } else {  // alpha at the sub-block end
            float tmp[8];
#pragma unroll
            for (int s = 0; s < 8; s++) tmp[s] = a[s];
            const int base = 8 * (NW - 1) * PP + t;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int idx = base + k * PP;
                alpha_step<LM>(tmp, sm.X[elem(idx)], __half2float(par[idx]));
            }
#pragma unroll
            for (int s = 0; s < 8; s++) sa[s] = tmp[s];
        }
        float spec[8];
#pragma unroll
        for (int s = 0; s < 8; s++) spec[s] = a[s];
#pragma unroll 1
        for (int w = NW - 1; w >= 0; w--) {
            const bool mid = (w > 0) && (w < NW - 1);
            float aw[8][8], u[8], v[8];
            int e[8];
            aw[0][0] = 0.f;
#pragma unroll
            for (int s = 1; s < 8; s++) aw[0][s] = mid ? sm.ckpt[((w - 1) * 7 + (s - 1)) * P + t] : spec[s];
            const int base = 8 * w * PP + t;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int idx = base + k * PP;
                e[k] = elem(idx);
                u[k] = sm.X[e[k]];
                v[k] = __half2float(par[idx]);
                if (k < 7) alpha_step_to<LM>(aw[k], u[k], v[k], aw[k + 1]);
            }
            norm8(b);
            unsigned acc = 0;
#pragma unroll
            for (int k = 7; k >= 0; k--) {
                const float ex = extrinsic<LM>(aw[k], b, v[k]);
                const float ec = fminf(fmaxf(ex, -c.clamp), c.clamp);
                const float es = c.q * ec;
                sm.X[e[k]] = __half2float(sm.sys[e[k]]) + es;
                if (want) {
                    const float lam = u[k] + ex;  // a-posteriori, :1038
                    if (!(lam < 0.f)) acc |= 1u << k;  // decision(): LLR < 0 -> 0 else 1, :862-879
                    if (lam < c.etT && lam > -c.etT) weak = true;
                    const int i = t * L + 8 * w + k;  // position in this SISO's order
                    if (g_llr) g_llr[i] = lam;
                    if (g_ext) g_ext[i] = es;
                }
                beta_step<LM>(b, u[k], v[k]);
            }
            if (w == w_sb) {
#pragma unroll
                for (int s = 0; s < 8; s++) sb[s] = b[s];
            }
            if (want) {
                changed |= acc ^ sm.dec[w * P + t];
                sm.dec[w * P + t] = (uint8_t)acc;
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
    float up[8], dn[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        up[s] = __shfl_up_sync(0xffffffffu, sa[s], 1);
        dn[s] = __shfl_down_sync(0xffffffffu, sb[s], 1);
    }
    if (lane == 31) {
#pragma unroll
        for (int s = 0; s < 8; s++) sm.edge[s * nwarps + warp] = sa[s];
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < 8; s++) sm.edge[(8 + s) * nwarps + warp] = sb[s];
    }
    __syncthreads();  // also orders this pass's X updates before the next pass's reads
    if (lane == 0 && warp > 0) {
#pragma unroll
        for (int s = 0; s < 8; s++) up[s] = sm.edge[s * nwarps + warp - 1];
    }
    if (lane == 31 && warp + 1 < nwarps) {
#pragma unroll
        for (int s = 0; s < 8; s++) dn[s] = sm.edge[(8 + s) * nwarps + warp + 1];
    }
#pragma unroll
    for (int s = 0; s < 8; s++) {
        if (!first_fixed) na[s] = up[s];
        if (!last_fixed) nb[s] = dn[s];
    }
    return changed;
}

template <int LLR_T, int LM>
__global__ void __launch_bounds__(256, 1) f32_kernel(F32Args A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FastGeom &g = A.g;
    const int P = g.P, PP = g.PP, NW = g.NW, L = g.L, K = g.K;
    const int W = L * PP, Wp = (W + 7) & ~7;
    Smem sm;
    sm.X = reinterpret_cast<float *>(smem_raw);
    sm.sys = reinterpret_cast<__half *>(sm.X + Wp);
    sm.par1 = sm.sys + Wp;
    sm.par2 = sm.par1 + Wp;
    sm.tab = reinterpret_cast<uint16_t *>(sm.par2 + Wp);
    sm.ckpt = reinterpret_cast<float *>(sm.tab + Wp);
    sm.edge = sm.ckpt + (size_t)g.n_ckpt * 7 * P;
    sm.dec = reinterpret_cast<uint8_t *>(sm.edge + 16 * (blockDim.x >> 5));
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int cb = blockIdx.x;
    const size_t row = (size_t)3 * K + 12;

    // ---- load + de-multiplex (demultiplex(), :1083-1127); element n = tt*L + j -> word j*PP + tt
    for (int n = tid; n < K; n += nthr) {
        const int tt = n / L, j = n - tt * L, ad = j * PP + tt;
        const __half s = __float2half_rn(load_llr<LLR_T>(A.llr, cb * row + 3 * (size_t)n));
        sm.sys[ad] = s;
        sm.X[ad] = __half2float(s);
        sm.par1[ad] = __float2half_rn(load_llr<LLR_T>(A.llr, cb * row + 3 * (size_t)n + 1));
        sm.par2[ad] = __float2half_rn(load_llr<LLR_T>(A.llr, cb * row + 3 * (size_t)n + 2));
    }
    for (int i = tid; i < W; i += nthr) sm.tab[i] = __ldg(A.tab2 + i);

    Pass c;
    c.P = P; c.PP = PP; c.NW = NW; c.G = g.G; c.L = L;
    c.q = A.ext_scale; c.clamp = A.ext_clamp; c.etT = A.et_threshold;
    const bool first_fixed = (tid == 0), last_fixed = (tid == P - 1);

    float na[2][8], nb[2][8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            na[s][j] = (first_fixed && j) ? kNegF : 0.f;  // known start state, :943-948
            nb[s][j] = 0.f;
        }
        if (last_fixed) {
            // termination folded into beta(K): three tail steps back from state 0, :950-954
            float bt[8];
#pragma unroll
            for (int j = 0; j < 8; j++) bt[j] = j ? kNegF : 0.f;
            for (int m = 2; m >= 0; m--) {
                const size_t o = cb * row + (size_t)3 * K + 6 * s + 2 * m;
                const float u = __half2float(__float2half_rn(load_llr<LLR_T>(A.llr, o)));
                const float v = __half2float(__float2half_rn(load_llr<LLR_T>(A.llr, o + 1)));
                beta_step<LM>(bt, u, v);
            }
            norm8(bt);
#pragma unroll
            for (int j = 0; j < 8; j++) nb[s][j] = bt[j];
        }
    }
    __syncthreads();

    const int T = K + kTail;
    float *g_llr = A.llr2 ? A.llr2 + (size_t)cb * T : nullptr;
    float *g_ext = A.ext2 ? A.ext2 + (size_t)cb * T : nullptr;
    int used = A.n_iter;
    for (int it = 0; it < A.n_iter; it++) {
        const bool last = (it == A.n_iter - 1);
        bool weak = false;
        siso_pass<LM, false>(c, sm, sm.par1, na[0], nb[0], first_fixed, last_fixed, false, nullptr, nullptr, weak);
        const bool want = A.early_term || last;
        const unsigned chg = siso_pass<LM, true>(c, sm, sm.par2, na[1], nb[1], first_fixed, last_fixed, want,
                                                 want ? g_llr : nullptr, want ? g_ext : nullptr, weak);
        if (A.early_term) {
            const int ch = __syncthreads_or((int)(chg != 0u || weak));
            if (it >= 1 && !ch) { used = it + 1; break; }
        }
    }

    // ---- hard decisions to natural order (random_deinterlvr_int, :1264): step i of SISO-2 is bit pi(i)
    if (A.bits && tid < P) {
        uint8_t *ob = A.bits + (size_t)cb * K;
        for (int w = 0; w < NW; w++) {
            const unsigned d8 = sm.dec[w * P + tid];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int e = sm.tab[(8 * w + k) * PP + tid];
                const int jj = e / PP, tt = e - jj * PP;
                ob[tt * L + jj] = (uint8_t)((d8 >> k) & 1u);
            }
        }
    }
    if (A.iters_used && tid == 0) A.iters_used[cb] = used;
    if (tid < kTail) {
        if (g_llr) g_llr[K + tid] = 0.f;
        if (g_ext) g_ext[K + tid] = 0.f;
    }
}

typedef void (*kernel_fn)(F32Args);
template <int LLR_T>
kernel_fn pick_t(int lm)
{
    return lm == 1 ? f32_kernel<LLR_T, 1> : (lm == 2 ? f32_kernel<LLR_T, 2> : f32_kernel<LLR_T, 0>);
}
kernel_fn pick(int llr_type, int lm)
{
    switch (llr_type) {
        case TDB200_LLR_F32: return pick_t<TDB200_LLR_F32>(lm);
        case TDB200_LLR_F64: return pick_t<TDB200_LLR_F64>(lm);
        case TDB200_LLR_F16: return pick_t<TDB200_LLR_F16>(lm);
        default: return pick_t<TDB200_LLR_S8>(lm);
    }
}

}  // namespace

int f32_smem_bytes(const FastGeom &g)
{
    const int nwarps = g.threads / 32;
    const int W = g.L * g.PP, Wp = (W + 7) & ~7;
    return 4 * Wp + 3 * 2 * Wp + 2 * Wp + 4 * g.n_ckpt * 7 * g.P + 4 * 16 * nwarps + ((g.NW * g.P + 15) & ~15);
}

cudaError_t f32_configure(const FastGeom &)
{
    int dev = 0, optin = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e0 != cudaSuccess) return e0;
    for (int t = TDB200_LLR_F64; t <= TDB200_LLR_F16; t++)
        for (int lm = 0; lm < 3; lm++) {
            cudaError_t e = cudaFuncSetAttribute(pick(t, lm), cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

cudaError_t launch_f32(const F32Args &a, cudaStream_t st, int *n_launches)
{
    pick(a.llr_type, a.logmap)<<<a.n_cb, a.g.threads, a.g.smem_bytes, st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

}  // namespace tdb200
