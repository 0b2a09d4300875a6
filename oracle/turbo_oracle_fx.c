/*
 * turbo_oracle_fx.c -- fixed-point sub-block-parallel max-log-MAP model.
 *
 * TEST INFRASTRUCTURE ONLY (see turbo_oracle.h).  This is NOT a restatement of reference
 * code: the reference's CPU path is the fp64 Log-MAP in turbo_oracle.c.  It is the bit-exact
 * integer specification of the throughput kernel TDB200_ALGO_MAXLOG_S16
 * (turbo_decoder_cuda_b200/csrc/tdb200_fast_kernel.cuh), written with plain int32 scalars and range
 * checks, so that the packed-s16x2 CUDA arithmetic can be verified bit for bit (hard
 * decisions AND extrinsics).  Its relation to the reference is algorithmic: it is
 * Log_MAP_decoder() (ITTC/log_map.cpp:898-1047) with
 *   - max* replaced by max (the reference's TYPE_DECODER 2 "MAX-LogMAP", log_map.h:26-29,
 *     which its CUDA prototypes implement: ITTC/CUDA/turboDecoderBianJieZhi.cu:205-401);
 *   - branch metrics shifted by the per-step constant (xs + xp + La/2) so that
 *     gamma(b,c) = b*U + c*V with U = Ls + La, V = Lp (full LLRs, 2^frac_bits fixed point);
 *   - the trellis cut into P = K/L sub-blocks that run concurrently, each started from the
 *     boundary metrics its neighbour produced in the previous iteration ("next-iteration
 *     initialisation", the boundary-value carry-over of turboDecoderBianJieZhi.cu:248,302-312);
 *   - the extrinsic scaled by 3/4 (turboDecoderBianJieZhi.cu:423-434 uses 0.77) and clamped;
 *   - tail bits folded into a fixed beta start vector (La is zero on the tail,
 *     log_map.cpp:1224-1227, so it never changes between iterations).
 * Agreement with the fp64 oracle is therefore statistical (BER/FER), tested separately.
 */
#include "turbo_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NS 8
#define FX_NEG (-14000) /* "impossible state" metric; see DESIGN.md range analysis */

static __thread int g_ovf;

static inline int chk(int v)
{
    if (v > 32767 || v < -32768) g_ovf = 1;
    return v;
}
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int add(int a, int b) { return chk(a + b); }

static inline int quant(float x, int frac_bits, int clip)
{
    float s = x * (float)(1 << frac_bits);
    int q;
    if (!(s == s)) return 0;       /* NaN -> erasure */
    if (s > 32767.0f) s = 32767.0f;
    if (s < -32767.0f) s = -32767.0f;
    q = (int)rintf(s);             /* round half to even == __float2int_rn */
    if (q > clip) q = clip;
    if (q < -clip) q = -clip;
    return q;
}

/* alpha' from alpha, with u = U, v = V, w = U+V */
static void alpha_step(const int *a, int u, int v, int *o)
{
    int w = add(u, v);
    o[0] = imax(a[0], add(a[1], w));
    o[4] = imax(add(a[0], w), a[1]);
    o[5] = imax(add(a[2], v), add(a[3], u));
    o[1] = imax(add(a[3], v), add(a[2], u));
    o[2] = imax(add(a[4], v), add(a[5], u));
    o[6] = imax(add(a[5], v), add(a[4], u));
    o[7] = imax(a[6], add(a[7], w));
    o[3] = imax(a[7], add(a[6], w));
}

/* beta (time i) from beta' (time i+1) */
static void beta_step(const int *b, int u, int v, int *o)
{
    int w = add(u, v);
    o[0] = imax(b[0], add(b[4], w));
    o[1] = imax(b[4], add(b[0], w));
    o[2] = imax(add(b[5], v), add(b[1], u));
    o[3] = imax(add(b[1], v), add(b[5], u));
    o[4] = imax(add(b[2], v), add(b[6], u));
    o[5] = imax(add(b[6], v), add(b[2], u));
    o[6] = imax(b[7], add(b[3], w));
    o[7] = imax(b[3], add(b[7], w));
}

/* extrinsic e = M1 - M0 (the common +U of the input-1 branches left out) */
static int extrinsic(const int *a, const int *b, int v)
{
    int m0a = imax(imax(add(a[0], b[0]), add(a[1], b[4])), imax(add(a[6], b[7]), add(a[7], b[3])));
    int m0b = imax(imax(add(a[2], b[5]), add(a[3], b[1])), imax(add(a[4], b[2]), add(a[5], b[6])));
    int m1a = imax(imax(add(a[0], b[4]), add(a[1], b[0])), imax(add(a[6], b[3]), add(a[7], b[7])));
    int m1b = imax(imax(add(a[2], b[1]), add(a[3], b[5])), imax(add(a[4], b[6]), add(a[5], b[2])));
    int m0 = imax(m0a, add(m0b, v));
    int m1 = imax(add(m1a, v), m1b);
    return chk(m1 - m0);
}


/* ------------------------------------------------------------------------------------------
 * Log-MAP in the same packed-16-bit arithmetic (TDB200_ALGO_LOGMAP_S16): max*(x,y) = max(x,y) +
 * c(|x-y|), the Jacobian logarithm the reference evaluates through E_algorithm()'s 16-step table
 * (ITTC/log_map.cpp:779-801, table :14-18), here as the linear fit c = max(0, T4 - |d|/4) in units of
 * 2^-frac_bits (T4 = 5 at 3 fractional bits: 0.625 - d/4; the least-squares line is 0.249(2.507 - d)).
 * The difference d never exists at full resolution: the two max* of a trellis butterfly share the
 * state-metric difference D = p - q, and d/4 is formed from floor((D-1)/4) and floor(g/4) (g = the
 * butterfly's branch-metric difference), which is what makes the correction cost one shift per
 * butterfly instead of one per max* on the device (there is no packed 16-bit shift instruction).
 * ------------------------------------------------------------------------------------------ */
static __thread int g_T4;      /* correction at d = 0, in fixed-point units */
static __thread int g_upper;   /* how the upper levels of the a-posteriori max* trees are corrected: 0 linear, 1 trapezoid, 2 not at all */
static __thread int g_T4L;     /* ... of the first level of the a-posteriori trees */
static __thread int g_exact;   /* exploration: 1 = linear correction on the exact difference, 2 = the exact Jacobian correction, rounded */
static __thread int g_F;
static __thread int g_uoff;    /* rounding offset of the generic max* */
static __thread int g_TT, g_TC;   /* trapezoid: c = min(TC, max(0, TT - |d|)) */

static inline int asr2(int x) { return x >> 2; } /* arithmetic shift: floor(x / 4) */
static inline int relu(int x) { return x > 0 ? x : 0; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* corrections of the two max* of a butterfly: ca for max*(p + g, q), cb for max*(q + g, p); g4 = floor(g / 4) */
static inline int corr_exact(int d)
{
    if (d < 0) d = -d;
    if (g_exact == 1) return relu((4 * g_T4 + 2 - d) >> 2);
    return (int)floor(log1p(exp(-(double)d / (double)(1 << g_F))) * (double)(1 << g_F) + 0.5);
}
static inline void bfly_corr_g(int p, int q, int g4, int g, int *ca, int *cb)
{
    if (g_exact) { *ca = corr_exact(p + g - q); *cb = corr_exact(q + g - p); return; }
    const int h = asr2(chk(p - q - 1)), nh = -h - 1;
    *ca = relu(imin(h + g4 + 1 + g_T4, nh - g4 + g_T4));
    *cb = relu(imin(nh + g4 + 1 + g_T4, h - g4 + g_T4));
}

static void alpha_step_lm(const int *a, int u, int v, int *o)
{
    const int w = add(u, v), w4 = asr2(w), gm4 = asr2(chk(u - v - 1));
    int ca, cb;
    bfly_corr_g(a[1], a[0], w4, w, &ca, &cb);
    o[0] = add(imax(add(a[1], w), a[0]), ca);
    o[4] = add(imax(add(a[0], w), a[1]), cb);
    bfly_corr_g(a[3], a[2], gm4, u - v, &ca, &cb);
    o[5] = add(imax(add(a[3], u), add(a[2], v)), ca);
    o[1] = add(imax(add(a[2], u), add(a[3], v)), cb);
    bfly_corr_g(a[5], a[4], gm4, u - v, &ca, &cb);
    o[2] = add(imax(add(a[5], u), add(a[4], v)), ca);
    o[6] = add(imax(add(a[4], u), add(a[5], v)), cb);
    bfly_corr_g(a[7], a[6], w4, w, &ca, &cb);
    o[7] = add(imax(add(a[7], w), a[6]), ca);
    o[3] = add(imax(add(a[6], w), a[7]), cb);
}

static void beta_step_lm(const int *b, int u, int v, int *o)
{
    const int w = add(u, v), w4 = asr2(w), gm4 = asr2(chk(u - v - 1));
    int ca, cb;
    bfly_corr_g(b[4], b[0], w4, w, &ca, &cb);
    o[0] = add(imax(add(b[4], w), b[0]), ca);
    o[1] = add(imax(add(b[0], w), b[4]), cb);
    bfly_corr_g(b[1], b[5], gm4, u - v, &ca, &cb);
    o[2] = add(imax(add(b[1], u), add(b[5], v)), ca);
    o[3] = add(imax(add(b[5], u), add(b[1], v)), cb);
    bfly_corr_g(b[6], b[2], gm4, u - v, &ca, &cb);
    o[4] = add(imax(add(b[6], u), add(b[2], v)), ca);
    o[5] = add(imax(add(b[2], u), add(b[6], v)), cb);
    bfly_corr_g(b[3], b[7], w4, w, &ca, &cb);
    o[6] = add(imax(add(b[3], w), b[7]), ca);
    o[7] = add(imax(add(b[7], w), b[3]), cb);
}

/* max* of two values that share nothing with anything else (the upper levels of the a-posteriori trees) */
static inline int maxstar_generic(int x, int y)
{
    const int mx = imax(x, y), mn = imin(x, y);
    const int e1 = chk(mn - mx - 1); /* -|d| - 1 */
    int c;
    if (g_exact) c = corr_exact(mx - mn);
    else if (g_upper == 2) c = 0;
    else if (g_upper == 1) c = imin(g_TC, relu(e1 + 1 + g_TT));
    else c = relu(asr2(e1) + g_uoff + g_T4);
    return add(mx, c);
}

/* first level of the a-posteriori trees: the four terms alpha_i + beta_m, alpha_j + beta_n (input 0) and
 * alpha_i + beta_n, alpha_j + beta_m (input 1) of a state pair (i, j) x (m, n); p = (aj, ai), q = (bn, bm)
 * oriented like the butterflies of the recursions, so hA and hB are the shifts those already formed */
static inline void lam_pair(int ai, int aj, int bm, int bn, int *same, int *cross)
{
    /* same: max*(aj + bn, ai + bm), d = (aj - ai) + (bn - bm);  cross: max*(aj + bm, ai + bn), d = (aj - ai) - (bn - bm) */
    const int hA = asr2(chk(aj - ai - 1)), nhA = -hA - 1;
    const int hB = asr2(chk(bn - bm - 1)), nhB = -hB - 1;
    int c1 = relu(imin(hA + hB + 1 + g_T4L, nhA + nhB + 1 + g_T4L));
    int c2 = relu(imin(hA + nhB + 1 + g_T4L, nhA + hB + 1 + g_T4L));
    if (g_exact) { c1 = corr_exact((aj + bn) - (ai + bm)); c2 = corr_exact((aj + bm) - (ai + bn)); }
    *same = add(imax(add(aj, bn), add(ai, bm)), c1);
    *cross = add(imax(add(aj, bm), add(ai, bn)), c2);
}

static int extrinsic_lm(const int *a, const int *b, int v)
{
    int s01, x01, s67, x67, s23, x23, s45, x45;
    lam_pair(a[0], a[1], b[0], b[4], &s01, &x01); /* input 0: a0+b0, a1+b4;  input 1: a1+b0, a0+b4 */
    lam_pair(a[6], a[7], b[7], b[3], &s67, &x67); /* input 0: a6+b7, a7+b3;  input 1: a7+b7, a6+b3 */
    lam_pair(a[2], a[3], b[5], b[1], &s23, &x23); /* input 0 (+v): a2+b5, a3+b1;  input 1: a3+b5, a2+b1 */
    lam_pair(a[4], a[5], b[2], b[6], &s45, &x45); /* input 0 (+v): a4+b2, a5+b6;  input 1: a5+b2, a4+b6 */
    const int m0a = maxstar_generic(s01, s67), m0b = maxstar_generic(s23, s45);
    const int m1a = maxstar_generic(x01, x67), m1b = maxstar_generic(x23, x45);
    const int m0 = maxstar_generic(m0a, add(m0b, v));
    const int m1 = maxstar_generic(add(m1a, v), m1b);
    return chk(m1 - m0);
}

static void normalise(int *m)
{
    int z = m[0];
    for (int s = 0; s < NS; s++) m[s] = chk(m[s] - z);
}

int tdo_fx_decode(const float *llr_in, const int *pi, const tdo_fx_params *p,
                  int *bits_out, int *le_out, int *overflow)
{
    const int K = p->K, L = p->sub_len, F = p->frac_bits;
    const int G = p->warmup;
    if (L < 8 || L % 8 || K % L || G < 0 || G % 8 || (G > L && G != 2 * L)) return -1;
    const int P = K / L;
    /* A guard longer than the sub-block (G = 2L, used with the 8-step sub-blocks of block sizes K = 8 x prime) spans
     * D = 2 sub-blocks: the warm-up of sub-block t starts from the START vector sub-block t-D had in the previous
     * iteration (its end vector for beta); where it would start before the first / after the last trellis step it
     * starts AT that step from the known vector instead. */
    const int D = G > L ? G / L : 1;
    g_ovf = 0;
    const int lm = p->logmap, lmw = lm && !p->lm_warm_maxlog;
    g_T4 = p->lm_t4 > 0 ? p->lm_t4 : 5 << (F > 3 ? F - 3 : 0);
    g_upper = p->lm_upper;
    g_TT = p->lm_tt; g_TC = p->lm_tc;
    g_uoff = p->lm_upper_off;
    g_exact = p->lm_exact; g_F = F;
    g_T4L = p->lm_t4_lam > 0 ? p->lm_t4_lam : g_T4;
#define ASTEP(lmf, a_, u_, v_, o_) ((lmf) ? alpha_step_lm(a_, u_, v_, o_) : alpha_step(a_, u_, v_, o_))
#define BSTEP(lmf, b_, u_, v_, o_) ((lmf) ? beta_step_lm(b_, u_, v_, o_) : beta_step(b_, u_, v_, o_))

    int *ys = (int *)malloc(sizeof(int) * K), *yp1 = (int *)malloc(sizeof(int) * K);
    int *yp2 = (int *)malloc(sizeof(int) * K), *X = (int *)malloc(sizeof(int) * K);
    int *alpha = (int *)malloc(sizeof(int) * NS * (L + 1));
    int(*niiA)[NS] = (int(*)[NS])calloc((size_t)2 * P * NS, sizeof(int));
    int(*niiB)[NS] = (int(*)[NS])calloc((size_t)2 * P * NS, sizeof(int));
    int(*newA)[NS] = (int(*)[NS])calloc((size_t)P * NS, sizeof(int));
    int(*newB)[NS] = (int(*)[NS])calloc((size_t)P * NS, sizeof(int));
    int *prev_bits = (int *)malloc(sizeof(int) * K);
    int *cur_bits = (int *)malloc(sizeof(int) * K);
    int *Xold = (int *)malloc(sizeof(int) * K);
    unsigned char *nat_bits = (unsigned char *)malloc(K);
    int crc_stop = 0;

    for (int i = 0; i < K; i++) {
        ys[i] = quant(llr_in[3 * i], F, p->llr_clip);
        yp1[i] = quant(llr_in[3 * i + 1], F, p->llr_clip);
        yp2[i] = quant(llr_in[3 * i + 2], F, p->llr_clip);
        X[i] = ys[i];
        prev_bits[i] = -1;
    }
    /* boundary vectors: known start state; tail folded into beta at step K */
    int fixedA[NS], tailB[2][NS];
    for (int j = 0; j < NS; j++) fixedA[j] = j ? FX_NEG : 0;
    for (int s = 0; s < 2; s++) {
        int b[NS], o[NS];
        for (int j = 0; j < NS; j++) {
            niiA[s * P + 0][j] = j ? FX_NEG : 0;
            b[j] = j ? FX_NEG : 0;
        }
        for (int m = 2; m >= 0; m--) {
            int u = quant(llr_in[3 * K + 6 * s + 2 * m], F, p->llr_clip);
            int v = quant(llr_in[3 * K + 6 * s + 2 * m + 1], F, p->llr_clip);
            BSTEP(lm, b, u, v, o);
            memcpy(b, o, sizeof(b));
        }
        normalise(b);
        memcpy(niiB[s * P + P - 1], b, sizeof(b));
        memcpy(tailB[s], b, sizeof(b));
    }

    const int et_T = p->et_threshold < 1 ? 1 : p->et_threshold;
    int it;
    for (it = 0; it < p->n_iter; it++) {
        int weak = 0;
        for (int s = 0; s < 2; s++) {
            const int *yp = s ? yp2 : yp1;
            /* warm-ups read the a-priori values as they were when the pass started (in the kernel
             * they run before the barrier that precedes the first in-place update of X) */
            memcpy(Xold, X, sizeof(int) * K);
            for (int t = 0; t < P; t++) {
                int b[NS], o[NS];
                /* ---- alpha warm-up over the last G steps of sub-block t-1, from the vector that
                 *      sub-block saved at its local step L-G in the previous iteration */
                int *a = alpha;
                memcpy(a, niiA[s * P + t], sizeof(int) * NS);
                int ka = -G;
                if (t > 0 && t * L + ka < 0) { /* (only with G > L) the warm-up would begin before step 0 */
                    ka = -t * L;
                    memcpy(a, fixedA, sizeof(fixedA));
                }
                if (t > 0)
                    for (int k = ka; k < 0; k++) {
                        int i = t * L + k, n = s ? pi[i] : i;
                        if ((k + G) % 8 == 0) normalise(a);
                        ASTEP(lmw, a, Xold[n], yp[i], o);
                        memcpy(a, o, sizeof(o));
                    }
                /* ---- beta warm-up over the first G steps of sub-block t+1 */
                memcpy(b, niiB[s * P + t], sizeof(b));
                int kb = G;
                if (t < P - 1 && (t + 1) * L + kb > K) { /* (only with G > L) ... would begin after the last step */
                    kb = K - (t + 1) * L;
                    memcpy(b, tailB[s], sizeof(b));
                }
                if (t < P - 1)
                    for (int k = kb - 1; k >= 0; k--) {
                        int i = (t + 1) * L + k, n = s ? pi[i] : i;
                        if (k % 8 == 7) normalise(b);
                        BSTEP(lmw, b, Xold[n], yp[i], o);
                        memcpy(b, o, sizeof(b));
                    }
                /* ---- forward: alpha at every step of the sub-block, normalised at window starts */
                for (int k = 0; k < L; k++) {
                    int i = t * L + k, n = s ? pi[i] : i;
                    if (k % 8 == 0) normalise(a + k * NS);
                    if (k == (G >= L ? 0 : L - G)) memcpy(newA[t], a + k * NS, sizeof(int) * NS);
                    ASTEP(lm, a + k * NS, X[n], yp[i], a + (k + 1) * NS);
                }
                if (G == 0) memcpy(newA[t], a + L * NS, sizeof(int) * NS);
                normalise(newA[t]);
                /* ---- backward: beta, extrinsic, a-posteriori, in-place update of X */
                if (G >= L) memcpy(newB[t], b, sizeof(b));
                for (int k = L - 1; k >= 0; k--) {
                    int i = t * L + k, n = s ? pi[i] : i;
                    if (k % 8 == 7) normalise(b);
                    int u = X[n], v = yp[i];
                    int e = lm ? extrinsic_lm(a + k * NS, b, v) : extrinsic(a + k * NS, b, v);
                    int lam = add(u, e);
                    int ec = e > p->ext_clip ? p->ext_clip : (e < -p->ext_clip - 1 ? -p->ext_clip - 1 : e); /* [-2^n, 2^n-1] */
                    int es = (p->ext_scale_q2 == 3) ? ((3 * ec) >> 2) : ec;
                    BSTEP(lm, b, u, v, o);
                    memcpy(b, o, sizeof(b));
                    X[n] = add(ys[n], es);
                    if (k == G) memcpy(newB[t], b, sizeof(b));
                    if (s == 0) nat_bits[i] = (unsigned char)(lam < 0 ? 0 : 1);
                    if (s == 1) {
                        cur_bits[n] = lam < 0 ? 0 : 1;
                        if (lam < et_T && lam > -et_T) weak = 1;
                        if (le_out) le_out[n] = es;
                    }
                }
                normalise(newB[t]);
            }
            /* synchronous hand-over of the boundary metrics to the neighbours (used next iteration) */
            for (int t = 0; t + D < P; t++) {
                memcpy(niiA[s * P + t + D], newA[t], sizeof(int) * NS);
                memcpy(niiB[s * P + t], newB[t + D], sizeof(int) * NS);
            }
            /* CRC stopping rule: after SISO-1 of the second and later iterations */
            if (s == 0 && p->early_term == 2 && it >= 1 && tdo_crc24(nat_bits, K, (unsigned)p->crc_poly) == 0) {
                for (int i = 0; i < K; i++) prev_bits[i] = nat_bits[i];
                crc_stop = 1;
                break;
            }
        }
        if (crc_stop) {
            it++;
            break;
        }
        int same = 1;
        for (int i = 0; i < K; i++) {
            if (cur_bits[i] != prev_bits[i]) same = 0;
            prev_bits[i] = cur_bits[i];
        }
        if (p->early_term == 1 && same && !weak && it >= 1) {
            it++;
            break;
        }
    }
    memcpy(bits_out, prev_bits, sizeof(int) * K);
    if (overflow) *overflow = g_ovf;
    free(ys); free(yp1); free(yp2); free(X); free(alpha);
    free(niiA); free(niiB); free(newA); free(newB); free(prev_bits); free(cur_bits); free(Xold); free(nat_bits);
    return it;
}
