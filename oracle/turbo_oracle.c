/*
 * turbo_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see turbo_oracle.h).
 *
 * Restates, in plain C and in the reference's floating-point operation order,
 * the decode path of xinxu27/turbo_decoder_cuda:
 *     TurboDecoding()    ITTC/log_map.cpp:1146-1280
 *     Log_MAP_decoder()  ITTC/log_map.cpp:898-1047
 *     E_algorithm()      ITTC/log_map.cpp:779-801 (+ LUT :14-18), E_algorithm_seq :817-829
 *     demultiplex()      ITTC/log_map.cpp:1083-1127
 *     (de)interleavers   ITTC/log_map.cpp:54-96, decision() :862-879
 *     gen_trellis()      ITTC/log_map.cpp:281-337, gen_qpp_index() :616-624
 * plus the encoder/channel needed to make test inputs (:451-583, main.cpp:174-202).
 * Arrays are time-major here (the reference is state-major); every add/sub/compare
 * is performed on the same operands in the same order, so results are expected to
 * agree with the compiled reference to the last bit except for the reference's
 * uninitialised tempmax[] (see tdo_siso).
 */
#include "turbo_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NS 8            /* states: 2^M_num_reg, ITTC/log_map.cpp:28-29 */
#define MREG 3          /* tail steps */
#define TDO_INFTY 1E20  /* ITTC/log_map.h:72-74 */

/* ------------------------------------------------------------------ trellis */

/* RSC with feedback 13_8 = 1011 and feed-forward 15_8 = 1101 (ITTC/log_map.h:34-36).
 * State integer = 4*s0 + 2*s1 + s2 with s0 the newest register (bin2int, :213-231). */
static void rsc_step(int state, int dk, int *next_state, int *parity)
{
    int s0 = (state >> 2) & 1, s1 = (state >> 1) & 1, s2 = state & 1;
    int ak = (dk + s1 + s2) & 1;          /* feedback taps g1 = 1,0,1,1  (:304-310) */
    *parity = (ak + s0 + s2) & 1;         /* forward taps  g2 = 1,1,0,1  (:253-258) */
    *next_state = (ak << 2) | (s0 << 1) | s1; /* shift register (:261-266) */
}

void tdo_gen_trellis(int *nextout, int *nextstat, int *lastout, int *laststat)
{
    for (int s = 0; s < NS; s++)
        for (int b = 0; b < 2; b++) {
            int ns, par;
            rsc_step(s, b, &ns, &par);
            nextout[s * 4 + 2 * b] = 2 * b - 1;
            nextout[s * 4 + 2 * b + 1] = 2 * par - 1;
            nextstat[s * 2 + b] = ns;
        }
    for (int s = 0; s < NS; s++)
        for (int b = 0; b < 2; b++) {
            int ns = nextstat[s * 2 + b];
            laststat[ns * 2 + b] = s;
            lastout[ns * 4 + 2 * b] = nextout[s * 4 + 2 * b];
            lastout[ns * 4 + 2 * b + 1] = nextout[s * 4 + 2 * b + 1];
        }
}

typedef struct {
    int nextout[NS * 4], nextstat[NS * 2], lastout[NS * 4], laststat[NS * 2];
} trellis_t;

static const trellis_t *trellis(void)
{
    static trellis_t t;
    static int ready = 0;
    if (!ready) {
        tdo_gen_trellis(t.nextout, t.nextstat, t.lastout, t.laststat);
        __sync_synchronize();
        ready = 1;
    }
    return &t;
}

/* ---------------------------------------------------------------- interleaver */

void tdo_qpp_index(int K, int f1, int f2, int *pi)
{
    /* staged modular form of ITTC/log_map.cpp:622, in 64-bit so any K is safe */
    for (int i = 0; i < K; i++) {
        long long a = ((long long)f2 * i) % K;
        pi[i] = (int)(((long long)f1 * i + (a * i) % K) % K);
    }
}

static const short lte_tab[188][3] = {
    {40,3,10},{48,7,12},{56,19,42},{64,7,16},{72,7,18},{80,11,20},{88,5,22},{96,11,24},{104,7,26},{112,41,84},
    {120,103,90},{128,15,32},{136,9,34},{144,17,108},{152,9,38},{160,21,120},{168,101,84},{176,21,44},{184,57,46},{192,23,48},
    {200,13,50},{208,27,52},{216,11,36},{224,27,56},{232,85,58},{240,29,60},{248,33,62},{256,15,32},{264,17,198},{272,33,68},
    {280,103,210},{288,19,36},{296,19,74},{304,37,76},{312,19,78},{320,21,120},{328,21,82},{336,115,84},{344,193,86},{352,21,44},
    {360,133,90},{368,81,46},{376,45,94},{384,23,48},{392,243,98},{400,151,40},{408,155,102},{416,25,52},{424,51,106},{432,47,72},
    {440,91,110},{448,29,168},{456,29,114},{464,247,58},{472,29,118},{480,89,180},{488,91,122},{496,157,62},{504,55,84},{512,31,64},
    {528,17,66},{544,35,68},{560,227,420},{576,65,96},{592,19,74},{608,37,76},{624,41,234},{640,39,80},{656,185,82},{672,43,252},
    {688,21,86},{704,155,44},{720,79,120},{736,139,92},{752,23,94},{768,217,48},{784,25,98},{800,17,80},{816,127,102},{832,25,52},
    {848,239,106},{864,17,48},{880,137,110},{896,215,112},{912,29,114},{928,15,58},{944,147,118},{960,29,60},{976,59,122},{992,65,124},
    {1008,55,84},{1024,31,64},{1056,17,66},{1088,171,204},{1120,67,140},{1152,35,72},{1184,19,74},{1216,39,76},{1248,19,78},{1280,199,240},
    {1312,21,82},{1344,211,252},{1376,21,86},{1408,43,88},{1440,149,60},{1472,45,92},{1504,49,846},{1536,71,48},{1568,13,28},{1600,17,80},
    {1632,25,102},{1664,183,104},{1696,55,954},{1728,127,96},{1760,27,110},{1792,29,112},{1824,29,114},{1856,57,116},{1888,45,354},{1920,31,120},
    {1952,59,610},{1984,185,124},{2016,113,420},{2048,31,64},{2112,17,66},{2176,171,136},{2240,209,420},{2304,253,216},{2368,367,444},{2432,265,456},
    {2496,181,468},{2560,39,80},{2624,27,164},{2688,127,504},{2752,143,172},{2816,43,88},{2880,29,300},{2944,45,92},{3008,157,188},{3072,47,96},
    {3136,13,28},{3200,111,240},{3264,443,204},{3328,51,104},{3392,51,212},{3456,451,192},{3520,257,220},{3584,57,336},{3648,313,228},{3712,271,232},
    {3776,179,236},{3840,331,120},{3904,363,244},{3968,375,248},{4032,127,168},{4096,31,64},{4160,33,130},{4224,43,264},{4288,33,134},{4352,477,408},
    {4416,35,138},{4480,233,280},{4544,357,142},{4608,337,480},{4672,37,146},{4736,71,444},{4800,71,120},{4864,37,152},{4928,39,462},{4992,127,234},
    {5056,39,158},{5120,39,80},{5184,31,96},{5248,113,902},{5312,41,166},{5376,251,336},{5440,43,170},{5504,21,86},{5568,43,174},{5632,45,176},
    {5696,45,178},{5760,161,120},{5824,89,182},{5888,323,184},{5952,47,186},{6016,23,94},{6080,47,190},{6144,263,480}};

int tdo_lte_num_sizes(void) { return 188; }
int tdo_lte_size_at(int idx) { return (idx >= 0 && idx < 188) ? lte_tab[idx][0] : -1; }

int tdo_lte_qpp_params(int K, int *f1, int *f2)
{
    for (int i = 0; i < 188; i++)
        if (lte_tab[i][0] == K) {
            *f1 = lte_tab[i][1];
            *f2 = lte_tab[i][2];
            return 0;
        }
    return -1;
}

/* ------------------------------------------------------------------- encoder */

/* rsc_encode(), ITTC/log_map.cpp:451-512: out[2i]=dk, out[2i+1]=parity, K+3 steps,
 * tail input chosen so the feedback sum is zero (:483-491). */
static void rsc_encode(const int *src, int K, int *out)
{
    int state = 0;
    for (int i = 0; i < K + MREG; i++) {
        int dk;
        if (i < K)
            dk = src[i];
        else {
            int s1 = (state >> 1) & 1, s2 = state & 1;
            dk = (s1 + s2) & 1;
        }
        int ns, par;
        rsc_step(state, dk, &ns, &par);
        out[2 * i] = dk;
        out[2 * i + 1] = par;
        state = ns;
    }
}

void tdo_turbo_encode(const int *bits, int K, const int *pi, int *coded)
{
    int T = K + MREG;
    int *rsc1 = (int *)malloc(sizeof(int) * 2 * T);
    int *rsc2 = (int *)malloc(sizeof(int) * 2 * T);
    int *in2 = (int *)calloc((size_t)K, sizeof(int));
    rsc_encode(bits, K, rsc1);
    for (int i = 0; i < K; i++) in2[i] = bits[pi[i]]; /* randominterleaver_int, :54-63 */
    rsc_encode(in2, K, rsc2);
    for (int i = 0; i < K; i++) { /* :566-571 */
        coded[3 * i] = rsc1[2 * i];
        coded[3 * i + 1] = rsc1[2 * i + 1];
        coded[3 * i + 2] = rsc2[2 * i + 1];
    }
    for (int i = 0; i < 2 * MREG; i++) { /* :574-578 */
        coded[3 * K + i] = rsc1[2 * K + i];
        coded[3 * K + 2 * MREG + i] = rsc2[2 * K + i];
    }
    free(rsc1);
    free(rsc2);
    free(in2);
}

/* ------------------------------------------------------------------- channel */

static uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

double tdo_sigma_from_ebn0(double ebn0_db, int K)
{
    /* ITTC/main.cpp:47,174 with MODULATION = 1: rate = K / (3K+12) */
    double rate = (double)K / (double)(3 * K + 4 * MREG);
    return pow(10.0, -ebn0_db / 20.0) * sqrt(0.5 / rate);
}

void tdo_channel_llr(const int *coded, int n, double sigma, unsigned long long seed,
                     unsigned long long stream, double *llr)
{
    const double two_pi = 6.283185307179586476925286766559;
    double kf2 = 2.0 / (sigma * sigma); /* demodule(): -Kf*(d1^2-d0^2) = 4*Kf*r, Kf = 1/(2 sigma^2) */
    for (int i = 0; i < n; i += 2) {
        uint64_t h = mix64(mix64(seed) ^ mix64(stream * 0x100000001B3ull + (uint64_t)(i >> 1)));
        uint64_t h2 = mix64(h);
        double u1 = ((double)(h >> 11) + 1.0) * (1.0 / 9007199254740992.0);
        double u2 = (double)(h2 >> 11) * (1.0 / 9007199254740992.0);
        double r = sqrt(-2.0 * log(u1));
        double n0 = r * cos(two_pi * u2), n1 = r * sin(two_pi * u2);
        llr[i] = kf2 * ((double)(2 * coded[i] - 1) + sigma * n0);
        if (i + 1 < n) llr[i + 1] = kf2 * ((double)(2 * coded[i + 1] - 1) + sigma * n1);
    }
}

/* ---------------------------------------------------------------------- max* */

static const double lut_index[16] = {0.0, 0.08824, 0.19587, 0.31026, 0.43275, 0.56508,
                                     0.70963, 0.86972, 1.0502, 1.2587, 1.5078, 1.8212,
                                     2.2522, 2.9706, 3.6764, 4.3758};
static const double lut_value[16] = {0.69315, 0.65, 0.6, 0.55, 0.5, 0.45, 0.4, 0.35,
                                     0.3, 0.25, 0.2, 0.15, 0.1, 0.05, 0.025, 0.0125};

double tdo_max_star_lut(double x, double y)
{
    double d = (y - x) > 0 ? (y - x) : (x - y);
    double c;
    if (d >= 4.3758)
        c = 0;
    else {
        int i;
        for (i = 0; i < 16 && d >= lut_index[i]; i++) {
        }
        c = lut_value[i - 1];
    }
    return (x > y ? x : y) + c;
}

static inline double max_only(double x, double y) { return x > y ? x : y; }

/* ---------------------------------------------------------------------- SISO */

void tdo_siso(const double *recs, const double *La, int terminated, double *LLR,
              int T, int algo, double tempmax_floor)
{
    const trellis_t *tr = trellis();
    double (*E)(double, double) = (algo == TDO_ALGO_MAXLOG) ? max_only : tdo_max_star_lut;

    double *alpha = (double *)malloc(sizeof(double) * NS * (T + 1)); /* [i][j] */
    double *beta = (double *)malloc(sizeof(double) * NS * (T + 1));
    double *g0 = (double *)malloc(sizeof(double) * NS * T); /* gamma, input 0, by from-state */
    double *g1 = (double *)malloc(sizeof(double) * NS * T);
    double *tmax = (double *)malloc(sizeof(double) * (T + 1));

    /* init, :943-960 */
    alpha[0] = 0;
    beta[T * NS] = 0;
    for (int j = 1; j < NS; j++) {
        alpha[j] = -TDO_INFTY;
        beta[T * NS + j] = terminated ? -TDO_INFTY : 0;
    }
    /* gamma, :962-972 */
    for (int i = 0; i < T; i++)
        for (int j = 0; j < NS; j++) {
            g0[i * NS + j] = -recs[2 * i] + recs[2 * i + 1] * tr->nextout[j * 4 + 1] - La[i] / 2;
            g1[i * NS + j] = recs[2 * i] + recs[2 * i + 1] * tr->nextout[j * 4 + 3] + La[i] / 2;
        }
    /* alpha forward, :975-1001 */
    for (int i = 1; i < T + 1; i++) {
        double *a = alpha + i * NS;
        const double *ap = alpha + (i - 1) * NS;
        for (int j = 0; j < NS; j++) {
            int l0 = tr->laststat[j * 2 + 0], l1 = tr->laststat[j * 2 + 1];
            double tx = g0[(i - 1) * NS + l0] + ap[l0];
            double ty = g1[(i - 1) * NS + l1] + ap[l1];
            a[j] = E(tx, ty);
        }
        /* :987-993 compares against uninitialised tempmax[i]; restated as a floor */
        double m;
        int j0 = 0;
        if (isnan(tempmax_floor)) {
            m = a[0];
            j0 = 1;
        } else
            m = tempmax_floor;
        for (int j = j0; j < NS; j++)
            if (m < a[j]) m = a[j];
        tmax[i] = m;
        for (int j = 0; j < NS; j++) a[j] = a[j] - m;
    }
    /* beta backward, :1004-1021 */
    for (int i = T - 1; i >= 0; i--) {
        double *b = beta + i * NS;
        const double *bn = beta + (i + 1) * NS;
        for (int j = 0; j < NS; j++) {
            double tx = g0[i * NS + j] + bn[tr->nextstat[j * 2 + 0]];
            double ty = g1[i * NS + j] + bn[tr->nextstat[j * 2 + 1]];
            b[j] = E(tx, ty);
        }
        for (int j = 0; j < NS; j++) b[j] = b[j] - tmax[i + 1];
    }
    /* LLR, :1024-1039; E_algorithm_seq folds j = 0..7 left to right (:817-829) */
    for (int i = 0; i < T; i++) {
        double t0[NS], t1[NS];
        for (int j = 0; j < NS; j++) {
            int l0 = tr->laststat[j * 2 + 0], l1 = tr->laststat[j * 2 + 1];
            t0[j] = g0[i * NS + l0] + alpha[i * NS + l0] + beta[(i + 1) * NS + j];
            t1[j] = g1[i * NS + l1] + alpha[i * NS + l1] + beta[(i + 1) * NS + j];
        }
        double m1 = E(t1[0], t1[1]), m0 = E(t0[0], t0[1]);
        for (int j = 2; j < NS; j++) {
            m1 = E(m1, t1[j]);
            m0 = E(m0, t0[j]);
        }
        LLR[i] = m1 - m0;
    }
    free(alpha);
    free(beta);
    free(g0);
    free(g1);
    free(tmax);
}

/* -------------------------------------------------------------------- decode */

void tdo_turbo_decode(const double *llr_in, int K, const int *pi, int n_iter, int algo,
                      int *bits_out, double *llr1_out, double *llr2_out, double *le_out)
{
    int T = K + MREG;
    double *yk = (double *)malloc(sizeof(double) * 4 * T);
    double *La = (double *)calloc(T, sizeof(double));
    double *Le = (double *)calloc(T, sizeof(double));
    double *LLR = (double *)calloc(T, sizeof(double));
    double *h = (double *)malloc(sizeof(double) * (3 * K + 4 * MREG));

    for (int i = 0; i < 3 * K + 4 * MREG; i++) h[i] = llr_in[i] * 0.5; /* :1202-1205 */

    /* demultiplex, :1083-1127 */
    for (int i = 0; i < K; i++) {
        yk[2 * i] = h[3 * i];
        yk[2 * i + 1] = h[3 * i + 1];
        yk[2 * T + 2 * i + 1] = h[3 * i + 2];
    }
    for (int i = 0; i < K; i++) yk[2 * T + 2 * i] = h[3 * pi[i]];
    for (int i = 0; i < 2 * MREG; i++) {
        yk[2 * K + i] = h[3 * K + i];
        yk[2 * T + 2 * K + i] = h[3 * K + 2 * MREG + i];
    }

    for (int it = 0; it < n_iter; it++) { /* :1217-1265 */
        for (int i = 0; i < K; i++) La[pi[i]] = Le[i]; /* random_deinterlvr_double, :87-96,1221 */
        for (int i = K; i < T; i++) La[i] = 0;
        /* tempmax floor 0.0: the reference on a clean (zero-filled) heap, see tdo_siso */
        tdo_siso(yk, La, 1, LLR, T, algo, 0.0);
        if (llr1_out && it == n_iter - 1) memcpy(llr1_out, LLR, sizeof(double) * T);
        for (int i = 0; i < T; i++) Le[i] = LLR[i] - La[i] - 2 * yk[2 * i]; /* :1234-1238 */
        for (int i = 0; i < K; i++) La[i] = Le[pi[i]]; /* randominterleaver_double, :76-85,1242 */
        for (int i = K; i < T; i++) La[i] = 0;
        tdo_siso(yk + 2 * T, La, 1, LLR, T, algo, 0.0);
        for (int i = 0; i < T; i++) Le[i] = LLR[i] - La[i] - 2 * yk[2 * T + 2 * i]; /* :1255-1259 */
        if (bits_out)
            for (int i = 0; i < K; i++) /* decision :862-879 + random_deinterlvr_int :1264 */
                bits_out[(size_t)K * it + pi[i]] = (LLR[i] < 0) ? 0 : 1;
    }
    if (llr2_out) memcpy(llr2_out, LLR, sizeof(double) * T);
    if (le_out) memcpy(le_out, Le, sizeof(double) * T);
    free(yk);
    free(La);
    free(Le);
    free(LLR);
    free(h);
}

typedef struct {
    const double *llr;
    int n_cb, K, n_iter, algo;
    const int *pi;
    int *bits_last;
    int next; /* shared work counter */
    pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *jb = (batch_job *)arg;
    int K = jb->K, L = 3 * K + 4 * MREG;
    int *bits = (int *)malloc(sizeof(int) * (size_t)K * jb->n_iter);
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        int c = jb->next++;
        pthread_mutex_unlock(&jb->mu);
        if (c >= jb->n_cb) break;
        tdo_turbo_decode(jb->llr + (size_t)c * L, K, jb->pi, jb->n_iter, jb->algo, bits, 0, 0, 0);
        if (jb->bits_last)
            memcpy(jb->bits_last + (size_t)c * K, bits + (size_t)K * (jb->n_iter - 1), sizeof(int) * K);
    }
    free(bits);
    return 0;
}

double tdo_turbo_decode_batch(const double *llr_in, int n_cb, int K, const int *pi,
                              int n_iter, int algo, int *bits_last, int n_threads)
{
    batch_job jb;
    jb.llr = llr_in; jb.n_cb = n_cb; jb.K = K; jb.n_iter = n_iter; jb.algo = algo;
    jb.pi = pi; jb.bits_last = bits_last; jb.next = 0;
    pthread_mutex_init(&jb.mu, 0);
    trellis(); /* build the table before threads start */
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], 0, batch_worker, &jb);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    pthread_mutex_destroy(&jb.mu);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
