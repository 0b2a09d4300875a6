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
    if (L < 8 || L % 8 || K % L || G < 0 || G % 8 || G > L) return -1;
    const int P = K / L;
    g_ovf = 0;

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
    for (int s = 0; s < 2; s++) {
        int b[NS], o[NS];
        for (int j = 0; j < NS; j++) {
            niiA[s * P + 0][j] = j ? FX_NEG : 0;
            b[j] = j ? FX_NEG : 0;
        }
        for (int m = 2; m >= 0; m--) {
            int u = quant(llr_in[3 * K + 6 * s + 2 * m], F, p->llr_clip);
            int v = quant(llr_in[3 * K + 6 * s + 2 * m + 1], F, p->llr_clip);
            beta_step(b, u, v, o);
            memcpy(b, o, sizeof(b));
        }
        normalise(b);
        memcpy(niiB[s * P + P - 1], b, sizeof(b));
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
                if (t > 0)
                    for (int k = -G; k < 0; k++) {
                        int i = t * L + k, n = s ? pi[i] : i;
                        if ((k + G) % 8 == 0) normalise(a);
                        alpha_step(a, Xold[n], yp[i], o);
                        memcpy(a, o, sizeof(o));
                    }
                /* ---- beta warm-up over the first G steps of sub-block t+1 */
                memcpy(b, niiB[s * P + t], sizeof(b));
                if (t < P - 1)
                    for (int k = G - 1; k >= 0; k--) {
                        int i = (t + 1) * L + k, n = s ? pi[i] : i;
                        if (k % 8 == 7) normalise(b);
                        beta_step(b, Xold[n], yp[i], o);
                        memcpy(b, o, sizeof(b));
                    }
                /* ---- forward: alpha at every step of the sub-block, normalised at window starts */
                for (int k = 0; k < L; k++) {
                    int i = t * L + k, n = s ? pi[i] : i;
                    if (k % 8 == 0) normalise(a + k * NS);
                    if (k == L - G) memcpy(newA[t], a + k * NS, sizeof(int) * NS);
                    alpha_step(a + k * NS, X[n], yp[i], a + (k + 1) * NS);
                }
                if (G == 0) memcpy(newA[t], a + L * NS, sizeof(int) * NS);
                normalise(newA[t]);
                /* ---- backward: beta, extrinsic, a-posteriori, in-place update of X */
                if (G == L) memcpy(newB[t], b, sizeof(b));
                for (int k = L - 1; k >= 0; k--) {
                    int i = t * L + k, n = s ? pi[i] : i;
                    if (k % 8 == 7) normalise(b);
                    int u = X[n], v = yp[i];
                    int e = extrinsic(a + k * NS, b, v);
                    int lam = add(u, e);
                    int ec = e > p->ext_clip ? p->ext_clip : (e < -p->ext_clip - 1 ? -p->ext_clip - 1 : e); /* [-2^n, 2^n-1] */
                    int es = (p->ext_scale_q2 == 3) ? ((3 * ec) >> 2) : ec;
                    beta_step(b, u, v, o);
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
            for (int t = 0; t + 1 < P; t++) {
                memcpy(niiA[s * P + t + 1], newA[t], sizeof(int) * NS);
                memcpy(niiB[s * P + t], newB[t + 1], sizeof(int) * NS);
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
