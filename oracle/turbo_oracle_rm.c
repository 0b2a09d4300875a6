/*
 * turbo_oracle_rm.c -- CPU oracle for the 3GPP wire format around the decode path:
 * TS 36.212 tail-bit multiplexing (5.1.3.2.2), sub-block interleaver and circular-buffer rate
 * matching (5.1.4.1), and the inverse (soft de-rate-matching with combining of repeated bits).
 * TEST INFRASTRUCTURE ONLY, like the rest of oracle/.
 *
 * The reference only DECLARES this stage -- void rate_match(int*, int, int*, int) and
 * void de_rate_match(double*, double*, int, int) in ITTC/main.h:23-24, call sites commented out at
 * ITTC/main.cpp:196,204 -- so there is no reference code to restate and no golden vector.
 *
 * Parity status: UNPINNED.  This file restates the published algorithm of TS 36.212 (V8+) section
 * 5.1.4.1 literally (matrix with <NULL> padding, column permutation, the pi(k) formula of the third
 * stream, the bit-selection loop with its k0), from memory: the container has no network and the
 * specification's text is not in /root/reference.  What the tests can and do check: structural
 * properties (bijection onto the non-NULL positions, N_D, K_pi, k0), agreement of the device's
 * table-driven implementation with this literal restatement for every block size, and
 * encode -> rate-match -> de-rate-match -> decode round trips.
 *
 * Input/output order on the turbo-code side is the REFERENCE's multiplex order
 * (ITTC/log_map.cpp:566-578): [3i]=x_i, [3i+1]=z_i, [3i+2]=z'_i, then (x,z)x3 of RSC1, (x',z')x3 of RSC2.
 */
#include <stdlib.h>
#include <string.h>

#include "turbo_oracle.h"

static const int kColPerm[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

/* 36.212 5.1.3.2.2: position in the reference's multiplex order of element k of stream s */
static int mux_pos(int K, int s, int k)
{
    if (k < K) return 3 * k + s;
    /* tails: x_{K+m} at 3K+2m, z_{K+m} at 3K+2m+1, x'_{K+m} at 3K+6+2m, z'_{K+m} at 3K+7+2m */
    const int X = 3 * K, Z = 3 * K + 1, Xp = 3 * K + 6, Zp = 3 * K + 7;
    static const int tab[3][4][2] = {
        /* d0: x_K, z_{K+1}, x'_K, z'_{K+1} */ {{0, 0}, {1, 1}, {2, 0}, {3, 1}},
        /* d1: z_K, x_{K+2}, z'_K, x'_{K+2} */ {{1, 0}, {0, 2}, {3, 0}, {2, 2}},
        /* d2: x_{K+1}, z_{K+2}, x'_{K+1}, z'_{K+2} */ {{0, 1}, {1, 2}, {2, 1}, {3, 2}}};
    const int kind = tab[s][k - K][0], m = tab[s][k - K][1];
    const int base = kind == 0 ? X : (kind == 1 ? Z : (kind == 2 ? Xp : Zp));
    return base + 2 * m;
}

int tdo_rm_geometry(int K, int *R, int *Kpi, int *ND)
{
    const int D = K + 4, r = (D + 31) / 32;
    if (R) *R = r;
    if (Kpi) *Kpi = 32 * r;
    if (ND) *ND = 32 * r - D;
    return 3 * 32 * r; /* K_w */
}

/* The circular buffer as a list of multiplex positions, -1 for <NULL>.  w[K_w].
 * F filler bits (5.1.2: the first F bits of a transport block's first code block): c_k = <NULL> for k < F, and
 * 5.1.3.2.1 sets d0_k = d1_k = <NULL> for those k (the encoder reads them as 0; d2 is transmitted). */
void tdo_rm_circular_buffer_f(int K, int F, int *w)
{
    int R, Kpi, ND;
    tdo_rm_geometry(K, &R, &Kpi, &ND);
    int *y = (int *)malloc(sizeof(int) * Kpi), *v = (int *)malloc(sizeof(int) * Kpi);
    for (int s = 0; s < 3; s++) {
        for (int k = 0; k < Kpi; k++) y[k] = (k < ND || (s < 2 && k - ND < F)) ? -1 : mux_pos(K, s, k - ND);
        if (s < 2) {
            /* rows of 32, columns permuted, read out column by column */
            for (int c = 0; c < 32; c++)
                for (int r = 0; r < R; r++) v[c * R + r] = y[r * 32 + kColPerm[c]];
        } else {
            for (int k = 0; k < Kpi; k++) v[k] = y[(kColPerm[k / R] + 32 * (k % R) + 1) % Kpi];
        }
        for (int k = 0; k < Kpi; k++) {
            if (s == 0) w[k] = v[k];
            else w[Kpi + 2 * k + (s - 1)] = v[k];
        }
    }
    free(y);
    free(v);
}

void tdo_rm_circular_buffer(int K, int *w) { tdo_rm_circular_buffer_f(K, 0, w); }

int tdo_rm_k0(int K, int rv, int Ncb)
{
    int R;
    const int Kw = tdo_rm_geometry(K, &R, NULL, NULL);
    if (Ncb <= 0 || Ncb > Kw) Ncb = Kw;
    return R * (2 * ((Ncb + 8 * R - 1) / (8 * R)) * rv + 2);
}

/* bit selection and pruning: sel[E] = multiplex position transmitted at e */
int tdo_rm_selection_f(int K, int E, int rv, int Ncb, int F, int *sel)
{
    int R;
    const int Kw = tdo_rm_geometry(K, &R, NULL, NULL);
    if (Ncb <= 0 || Ncb > Kw) Ncb = Kw;
    int *w = (int *)malloc(sizeof(int) * Kw);
    tdo_rm_circular_buffer_f(K, F, w);
    int any = 0;
    for (int k = 0; k < Ncb; k++) any |= (w[k] >= 0);
    if (!any) { free(w); return -1; }
    const int k0 = tdo_rm_k0(K, rv, Ncb);
    int k = 0;
    for (long j = 0; k < E; j++) {
        const int p = w[(k0 + j) % Ncb];
        if (p >= 0) sel[k++] = p;
    }
    free(w);
    return 0;
}

int tdo_rm_selection(int K, int E, int rv, int Ncb, int *sel) { return tdo_rm_selection_f(K, E, rv, Ncb, 0, sel); }

int tdo_rate_match_f(const int *coded, int K, int E, int rv, int Ncb, int F, int *e_bits)
{
    int *sel = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    if (tdo_rm_selection_f(K, E, rv, Ncb, F, sel)) { free(sel); return -1; }
    for (int e = 0; e < E; e++) e_bits[e] = coded[sel[e]];
    free(sel);
    return 0;
}
int tdo_rate_match(const int *coded, int K, int E, int rv, int Ncb, int *e_bits) { return tdo_rate_match_f(coded, K, E, rv, Ncb, 0, e_bits); }

/* soft inverse with filler bits: the receiver knows the F filler bits and their parity-1 bits are 0 (the encoder starts in
 * state 0 and stays there while it reads zeros), so those positions -- never transmitted -- come out as the fixed value
 * `fill` (a confident "0": negative in this library's sign convention); with accumulate they are set, not added */
int tdo_rate_dematch_f(const double *e_llr, int K, int E, int rv, int Ncb, int F, int accumulate, double fill, double *llr)
{
    int *sel = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    if (tdo_rm_selection_f(K, E, rv, Ncb, F, sel)) { free(sel); return -1; }
    const int NL = 3 * K + 12;
    double *sum = (double *)calloc(NL, sizeof(double));
    for (int e = 0; e < E; e++) sum[sel[e]] += e_llr[e];
    for (int n = 0; n < NL; n++) llr[n] = accumulate ? llr[n] + sum[n] : sum[n];
    for (int k = 0; k < F; k++) llr[3 * k] = llr[3 * k + 1] = fill;
    free(sum);
    free(sel);
    return 0;
}

/* soft inverse: repeated positions are summed in transmission order, punctured ones stay 0
 * (accumulate != 0: new = old + this transmission's sums -- HARQ combining of retransmissions) */
int tdo_rate_dematch(const double *e_llr, int K, int E, int rv, int Ncb, int accumulate, double *llr)
{
    int *sel = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    if (tdo_rm_selection(K, E, rv, Ncb, sel)) { free(sel); return -1; }
    const int NL = 3 * K + 12;
    double *sum = (double *)calloc(NL, sizeof(double));
    for (int e = 0; e < E; e++) sum[sel[e]] += e_llr[e];
    for (int n = 0; n < NL; n++) llr[n] = accumulate ? llr[n] + sum[n] : sum[n];
    free(sum);
    free(sel);
    return 0;
}

/* the same in float (the device's accumulation type for float / half inputs) */
int tdo_rate_dematch_f32(const float *e_llr, int K, int E, int rv, int Ncb, int accumulate, float *llr)
{
    int *sel = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    if (tdo_rm_selection(K, E, rv, Ncb, sel)) { free(sel); return -1; }
    const int NL = 3 * K + 12;
    float *sum = (float *)calloc(NL, sizeof(float));
    for (int e = 0; e < E; e++) sum[sel[e]] += e_llr[e];
    for (int n = 0; n < NL; n++) llr[n] = accumulate ? llr[n] + sum[n] : sum[n];
    free(sum);
    free(sel);
    return 0;
}
