/*
 * turbo_oracle_mod.c -- CPU oracle for the mapper / soft demapper either side of the decode path
 * (SURVEY.md 8f.4).  TEST INFRASTRUCTURE ONLY, like the rest of oracle/.
 *
 * Restates ITTC/modanddem.cpp:
 *   module()     :175-186  (bit groups -> constellation points; tables :7-71)
 *   demodule()   :674-686  (max-log demapper: LLR_b = -Kf * (min_{j: bit b = 1} d_j - min_{j: bit b = 0} d_j),
 *                           d_j = squared distance to point j, :73-86; one exhaustive scan per bit, :189-672)
 * and adds the fp32 model of the device demapper that feeds the throughput decoder
 * (turbo_decoder_cuda_b200/csrc/tdb200_modem.cu): same metric, evaluated per axis where the
 * constellation is a product of two one-dimensional level sets (BPSK, QPSK, 16QAM, 64QAM).
 *
 * Parity status: PINNED -- tests/test_oracle.py checks tdo_modulate / tdo_demap_f64 against the
 * reference's own module()/demodule() compiled in place (oracle/_ref) and against
 * tests/golden/modem_golden.npz generated from that build.
 */
#include <math.h>
#include <stdint.h>

#include "turbo_oracle.h"

/* ---- constellations.  The point with index j (bits MSB first, except 8PSK whose index is built
 *      LSB first, modanddem.cpp:136) is (lvI[j >> nq], lvQ[j & (2^nq - 1)]) for the product
 *      constellations, (psk_i[j], psk_q[j]) for 8PSK. */
static const double lv_bpsk[2] = {-1.0, 1.0};                                                  /* :7-10 */
static const double lv_qpsk[2] = {0.7071, -0.7071};                                            /* :17-24 */
static const double lv_16[4] = {-0.948683, -0.316228, 0.948683, 0.316228};                     /* :35-49 */
static const double lv_64[8] = {0.4629, 0.1543, 0.7615, 1.0801, -0.4629, -0.1543, -0.7615, -1.0801}; /* :51-71 */
static const double psk_i[8] = {-0.7071, -1, 0, 0.7071, 0, -0.7071, 0.7071, 1};                /* :26-29 */
static const double psk_q[8] = {0.7071, 0, 1, 0.7071, -1, -0.7071, -0.7071, 0};                /* :31-34 */

/* number of bits carried by the I axis / the Q axis; 8PSK is not a product */
static int axis_bits(int M, int *nq)
{
    switch (M) {
        case 1: *nq = 0; return 1;
        case 2: *nq = 1; return 1;
        case 4: *nq = 2; return 2;
        case 6: *nq = 3; return 3;
        default: *nq = 0; return 0;
    }
}
static const double *levels(int M)
{
    return M == 1 ? lv_bpsk : (M == 2 ? lv_qpsk : (M == 4 ? lv_16 : lv_64));
}

int tdo_mod_point(int M, int j, double *pi, double *pq)
{
    if (M == 3) { *pi = psk_i[j & 7]; *pq = psk_q[j & 7]; return 0; }
    int nq, ni = axis_bits(M, &nq);
    if (!ni) return -1;
    const double *lv = levels(M);
    *pi = lv[j >> nq];
    *pq = nq ? lv[j & ((1 << nq) - 1)] : 0.0;
    return 0;
}

/* module(): n_bits must be a multiple of M */
int tdo_modulate(const int *bits, int n_bits, int M, double *si, double *sq)
{
    if (!(M == 1 || M == 2 || M == 3 || M == 4 || M == 6) || n_bits % M) return -1;
    for (int s = 0; s < n_bits / M; s++) {
        int j = 0;
        if (M == 3) j = bits[3 * s + 2] * 4 + bits[3 * s + 1] * 2 + bits[3 * s];
        else for (int b = 0; b < M; b++) j = 2 * j + bits[M * s + b];
        tdo_mod_point(M, j, &si[s], &sq[s]);
    }
    return 0;
}

/* demodule() in the reference's own order of operations: per bit an ascending scan over all
 * 2^M points, strict '<' updates, then -Kf * (min1 - min0).  out[M*s + b] belongs to index bit
 * (M-1-b) (MSB first), for 8PSK to index bit b (:329,355,380). */
int tdo_demap_f64(const double *si, const double *sq, int n_sym, int M, double kf, double *out)
{
    if (!(M == 1 || M == 2 || M == 3 || M == 4 || M == 6)) return -1;
    const int np = 1 << M;
    for (int s = 0; s < n_sym; s++) {
        for (int b = 0; b < M; b++) {
            const int mask = (M == 3) ? (1 << b) : (1 << (M - 1 - b));
            double m1 = (M <= 2) ? (double)0x7fffffffffff : (double)0x7fffffff, m0 = m1; /* :198,237 vs :304,394,510 */
            for (int j = 0; j < np; j++) {
                double pi, pq;
                tdo_mod_point(M, j, &pi, &pq);
                const double dr = si[s] - pi, di = sq[s] - pq;
                const double d = dr * dr + di * di;
                if (j & mask) { if (d < m1) m1 = d; }
                else          { if (d < m0) m0 = d; }
            }
            out[M * s + b] = -kf * (m1 - m0);
        }
    }
    return 0;
}

/* ---- fp32 model of the device demapper (bit-exact mirror; built with -ffp-contract=off) */
static float sqf(float a) { return a * a; }

static void axis_f32(float v, const double *lv, int nb, float kf, float *out)
{
    float d[8];
    const int n = 1 << nb;
    for (int l = 0; l < n; l++) d[l] = sqf(v - (float)lv[l]);
    for (int b = 0; b < nb; b++) {
        const int mask = n >> (b + 1);
        float m1 = 0, m0 = 0;
        int h1 = 0, h0 = 0;
        for (int l = 0; l < n; l++) {
            if (l & mask) { m1 = h1 ? fminf(m1, d[l]) : d[l]; h1 = 1; }
            else          { m0 = h0 ? fminf(m0, d[l]) : d[l]; h0 = 1; }
        }
        out[b] = -kf * (m1 - m0);
    }
}

int tdo_demap_f32(const float *si, const float *sq, int n_sym, int M, float kf, float *out)
{
    if (!(M == 1 || M == 2 || M == 3 || M == 4 || M == 6)) return -1;
    for (int s = 0; s < n_sym; s++) {
        if (M == 3) {
            float d[8];
            for (int j = 0; j < 8; j++) d[j] = sqf(si[s] - (float)psk_i[j]) + sqf(sq[s] - (float)psk_q[j]);
            for (int b = 0; b < 3; b++) {
                float m1 = 0, m0 = 0;
                int h1 = 0, h0 = 0;
                for (int j = 0; j < 8; j++) {
                    if (j & (1 << b)) { m1 = h1 ? fminf(m1, d[j]) : d[j]; h1 = 1; }
                    else              { m0 = h0 ? fminf(m0, d[j]) : d[j]; h0 = 1; }
                }
                out[3 * s + b] = -kf * (m1 - m0);
            }
        } else {
            int nq, ni = axis_bits(M, &nq);
            axis_f32(si[s], levels(M), ni, kf, out + M * s);
            if (nq) axis_f32(sq[s], levels(M), nq, kf, out + M * s + ni);
        }
    }
    return 0;
}

/* the throughput decoder's channel-value quantiser (turbo_oracle_fx.c: quant()) applied to fp32
 * demapper outputs: the 8-bit values the device demapper hands to the s16 kernel */
void tdo_quant_s8(const float *llr, int n, int frac_bits, int clip, int8_t *out)
{
    const float scale = (float)(1 << frac_bits);
    for (int i = 0; i < n; i++) {
        float s = llr[i] * scale;
        int q = 0;
        if (s == s) {
            s = fminf(fmaxf(s, -32767.0f), 32767.0f);
            q = (int)lrintf(s);
            q = q > clip ? clip : (q < -clip ? -clip : q);
        }
        out[i] = (int8_t)q;
    }
}
