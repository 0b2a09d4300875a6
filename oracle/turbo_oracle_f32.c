/*
 * turbo_oracle_f32.c -- fp32 sub-block-parallel (windowed) Log-MAP / max-log-MAP model.
 *
 * TEST INFRASTRUCTURE ONLY (see turbo_oracle.h).  Like turbo_oracle_fx.c this is NOT a
 * restatement of reference code: it is the plain-C specification of the CUDA kernels
 * TDB200_ALGO_LOGMAP_F32 / TDB200_ALGO_MAXLOG_F32 (turbo_decoder_cuda_b200/csrc/tdb200_f32.cu):
 * the decode path of TurboDecoding() / Log_MAP_decoder() (ITTC/log_map.cpp:1146-1280, :898-1047)
 *   - in single precision, with channel values rounded to fp16 (the kernel's shared-memory format);
 *   - max*(x,y) = max(x,y) + ln(1 + e^-|x-y|) evaluated with exp2f/log2f (the reference's
 *     E_algorithm, :779-801, tabulates the same correction in 16 steps), or plain max;
 *   - branch metrics gamma(b,c) = b*U + c*V, U = Ls + La, V = Lp (full LLRs);
 *   - the trellis cut into P = K/L sub-blocks with next-iteration initialisation + G warm-up steps,
 *     alpha/beta normalised every 8 steps -- the schedule of turbo_oracle_fx.c.
 * Every float operation is written in the order the kernel performs it (the kernel is compiled
 * with -fmad=false, this file with -ffp-contract=off), so the max-log variant is bit-exact and the
 * Log-MAP variant differs only by the hardware ex2/lg2 approximations.
 */
#include "turbo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NS 8
#define F32_NEG (-1.0e9f)

static int g_logmap;

static inline float mx(float x, float y)
{
    float m = x > y ? x : y;
    if (!g_logmap) return m;
    if (g_logmap == 2) return m + fmaxf(fmaf(-0.24904f, fabsf(x - y), 0.62429345f), 0.0f); /* linear correction, fused like the kernel's FFMA */
    float d = fabsf(x - y) * -1.4426950408889634f;
    float e = exp2f(d);
    float l = log2f(1.0f + e);
    return m + l * 0.6931471805599453f;
}

static inline float h16(float x)
{
    if (!(x == x)) x = 0.0f; /* NaN -> erasure */
    if (x > 65504.0f) x = 65504.0f; /* the finite range of binary16: no infinities in the recursions */
    if (x < -65504.0f) x = -65504.0f;
    return (float)(_Float16)x; /* round to nearest even, like __float2half_rn */
}

static void alpha_step(const float *a, float u, float v, float *o)
{
    float w = u + v;
    float o0 = mx(a[0], a[1] + w), o4 = mx(a[0] + w, a[1]);
    float o5 = mx(a[2] + v, a[3] + u), o1 = mx(a[3] + v, a[2] + u);
    float o2 = mx(a[4] + v, a[5] + u), o6 = mx(a[5] + v, a[4] + u);
    float o7 = mx(a[6], a[7] + w), o3 = mx(a[7], a[6] + w);
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5; o[6] = o6; o[7] = o7;
}

static void beta_step(const float *b, float u, float v, float *o)
{
    float w = u + v;
    float o0 = mx(b[0], b[4] + w), o1 = mx(b[4], b[0] + w);
    float o2 = mx(b[5] + v, b[1] + u), o3 = mx(b[1] + v, b[5] + u);
    float o4 = mx(b[2] + v, b[6] + u), o5 = mx(b[6] + v, b[2] + u);
    float o6 = mx(b[7], b[3] + w), o7 = mx(b[3], b[7] + w);
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5; o[6] = o6; o[7] = o7;
}

static float extrinsic(const float *a, const float *b, float v)
{
    float m0a = mx(mx(a[0] + b[0], a[1] + b[4]), mx(a[6] + b[7], a[7] + b[3]));
    float m0b = mx(mx(a[2] + b[5], a[3] + b[1]), mx(a[4] + b[2], a[5] + b[6]));
    float m1a = mx(mx(a[0] + b[4], a[1] + b[0]), mx(a[6] + b[3], a[7] + b[7]));
    float m1b = mx(mx(a[2] + b[1], a[3] + b[5]), mx(a[4] + b[6], a[5] + b[2]));
    return mx(m1a + v, m1b) - mx(m0a, m0b + v);
}

static void normalise(float *m)
{
    float z = m[0];
    m[0] = 0.0f;
    for (int s = 1; s < NS; s++) m[s] -= z;
}

int tdo_f32_decode(const float *llr_in, const int *pi, const tdo_f32_params *p,
                   int *bits_out, float *llr_out, float *le_out)
{
    const int K = p->K, L = p->sub_len, G = p->warmup;
    if (L < 8 || L % 8 || K % L || G < 0 || G % 8 || G > L) return -1;
    const int P = K / L;
    g_logmap = p->logmap;

    float *ys = malloc(sizeof(float) * K), *yp1 = malloc(sizeof(float) * K), *yp2 = malloc(sizeof(float) * K);
    float *X = malloc(sizeof(float) * K), *Xold = malloc(sizeof(float) * K);
    float *alpha = malloc(sizeof(float) * NS * (L + 1));
    float(*niiA)[NS] = calloc((size_t)2 * P * NS, sizeof(float));
    float(*niiB)[NS] = calloc((size_t)2 * P * NS, sizeof(float));
    float(*newA)[NS] = calloc((size_t)P * NS, sizeof(float));
    float(*newB)[NS] = calloc((size_t)P * NS, sizeof(float));
    int *prev_bits = malloc(sizeof(int) * K), *cur_bits = malloc(sizeof(int) * K);

    for (int i = 0; i < K; i++) {
        ys[i] = h16(llr_in[3 * i]);
        yp1[i] = h16(llr_in[3 * i + 1]);
        yp2[i] = h16(llr_in[3 * i + 2]);
        X[i] = ys[i];
        prev_bits[i] = -1;
    }
    for (int s = 0; s < 2; s++) {
        float b[NS], o[NS];
        for (int j = 0; j < NS; j++) {
            niiA[s * P + 0][j] = j ? F32_NEG : 0.0f;
            b[j] = j ? F32_NEG : 0.0f;
        }
        for (int m = 2; m >= 0; m--) {
            float u = h16(llr_in[3 * K + 6 * s + 2 * m]);
            float v = h16(llr_in[3 * K + 6 * s + 2 * m + 1]);
            beta_step(b, u, v, o);
            memcpy(b, o, sizeof(b));
        }
        normalise(b);
        memcpy(niiB[s * P + P - 1], b, sizeof(b));
    }

    int it;
    for (it = 0; it < p->n_iter; it++) {
        int weak = 0;
        for (int s = 0; s < 2; s++) {
            const float *yp = s ? yp2 : yp1;
            memcpy(Xold, X, sizeof(float) * K);
            for (int t = 0; t < P; t++) {
                float b[NS], o[NS];
                float *a = alpha;
                memcpy(a, niiA[s * P + t], sizeof(float) * NS);
                if (t > 0)
                    for (int k = -G; k < 0; k++) {
                        int i = t * L + k, n = s ? pi[i] : i;
                        if ((k + G) % 8 == 0) normalise(a);
                        alpha_step(a, Xold[n], yp[i], o);
                        memcpy(a, o, sizeof(o));
                    }
                memcpy(b, niiB[s * P + t], sizeof(b));
                if (t < P - 1)
                    for (int k = G - 1; k >= 0; k--) {
                        int i = (t + 1) * L + k, n = s ? pi[i] : i;
                        if (k % 8 == 7) normalise(b);
                        beta_step(b, Xold[n], yp[i], o);
                        memcpy(b, o, sizeof(b));
                    }
                for (int k = 0; k < L; k++) {
                    int i = t * L + k, n = s ? pi[i] : i;
                    if (k % 8 == 0) normalise(a + k * NS);
                    if (k == L - G) memcpy(newA[t], a + k * NS, sizeof(float) * NS);
                    alpha_step(a + k * NS, X[n], yp[i], a + (k + 1) * NS);
                }
                if (G == 0) memcpy(newA[t], a + L * NS, sizeof(float) * NS);
                normalise(newA[t]);
                if (G == L) memcpy(newB[t], b, sizeof(b));
                for (int k = L - 1; k >= 0; k--) {
                    int i = t * L + k, n = s ? pi[i] : i;
                    if (k % 8 == 7) normalise(b);
                    float u = X[n], v = yp[i];
                    float e = extrinsic(a + k * NS, b, v);
                    float lam = u + e;
                    float ec = fminf(fmaxf(e, -p->ext_clamp), p->ext_clamp);
                    float es = p->ext_scale * ec;
                    beta_step(b, u, v, o);
                    memcpy(b, o, sizeof(b));
                    X[n] = ys[n] + es;
                    if (k == G) memcpy(newB[t], b, sizeof(b));
                    if (s == 1) {
                        cur_bits[n] = lam < 0.0f ? 0 : 1;
                        if (lam < p->et_threshold && lam > -p->et_threshold) weak = 1;
                        if (llr_out) llr_out[i] = lam; /* SISO-2 (interleaved) order */
                        if (le_out) le_out[i] = es;
                    }
                }
                normalise(newB[t]);
            }
            for (int t = 0; t + 1 < P; t++) {
                memcpy(niiA[s * P + t + 1], newA[t], sizeof(float) * NS);
                memcpy(niiB[s * P + t], newB[t + 1], sizeof(float) * NS);
            }
        }
        int same = 1;
        for (int i = 0; i < K; i++) {
            if (cur_bits[i] != prev_bits[i]) same = 0;
            prev_bits[i] = cur_bits[i];
        }
        if (p->early_term && same && !weak && it >= 1) {
            it++;
            break;
        }
    }
    memcpy(bits_out, prev_bits, sizeof(int) * K);
    free(ys); free(yp1); free(yp2); free(X); free(Xold); free(alpha);
    free(niiA); free(niiB); free(newA); free(newB); free(prev_bits); free(cur_bits);
    return it;
}
