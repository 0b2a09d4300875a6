/*
 * ref_harness.cpp -- C-ABI shim around the UNMODIFIED reference decoder.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked (by oracle/Makefile) with the reference's own
 * translation units compiled where they lie:
 *     /root/reference/ITTC/log_map.cpp  /root/reference/ITTC/modanddem.cpp
 * into oracle/_ref/libittc_ref.so.  No reference source is copied into this repo.
 * This TU plays the role of ITTC/main.cpp: it includes ITTC/main.h, which DEFINES the
 * configuration globals (source_length, f1, f2, ... -- main.h:6-11), and calls the
 * reference entry points declared there (main.h:13-24).
 */
#include "main.h" /* -I/root/reference/ITTC */

#include <malloc.h>

#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

/* External-linkage symbols of ITTC/log_map.cpp that main.h does not declare. */
typedef struct {
    int *mx_nextout, *mx_nextstat, *mx_lastout, *mx_laststat; /* ITTC/log_map.h:61-68 */
} REF_TRELLIS;
extern REF_TRELLIS turbo_trellis;  /* ITTC/log_map.h:85 (same layout as TURBO_TRELLIS) */
extern int *index_randomintlvr;    /* ITTC/log_map.h:79 */
extern int M_num_reg;              /* ITTC/log_map.cpp:28 */
void Log_MAP_decoder(double *recs, double *La, int terminated, double *LLR, int len_total); /* :898 */
void demultiplex(double *rec, int len_info, double *yk);                                    /* :1083 */
void randominterleaver_double(double *in, double *out, int *index, int length);             /* :76 */
void random_deinterlvr_double(double *out, double *in, int *index, int length);             /* :87 */
void random_deinterlvr_int(int *out, int *in, int *index, int length);                      /* :65 */
void decision(double *LLR, int length, int *output);                                        /* :862 */
double E_algorithm(double x, double y);                                                     /* :779 */

static bool g_ready = false;

extern "C" {

int ref_n_iteration_macro(void) { return 15; } /* N_ITERATION, ITTC/log_map.h:30 */

/* The reference reads tempmax[] before writing it (log_map.cpp:925,989).  In its own process the
 * heap is clean and the value is 0 or an earlier, benign tempmax; inside a Python process the
 * recycled chunk can hold anything (1e300, NaN) and wreck a decode.  M_PERTURB = 0xFF makes glibc
 * fill every malloc'd block with 0x00, i.e. the reference behaves as on a fresh heap:
 * tempmax[i] = max(0, max_j alpha_j).  No reference source is touched. */
void ref_heap_zeroing(int on) { mallopt(M_PERTURB, on ? 0xFF : 0); }

/* M_PERTURB does not reach glibc's per-thread cache (blocks <= ~1 KB, i.e. tempmax[] for K <= 125):
 * park zero-filled blocks of that size in the cache so the next malloc returns one of them. */
static void scrub_small_blocks(size_t bytes)
{
    if (bytes > 1040) return;
    void *p[8];
    for (int i = 0; i < 8; i++) p[i] = calloc(1, bytes);
    for (int i = 0; i < 8; i++) free(p[i]);
}

/* main.cpp:29-37,106 */
void ref_init(int K, int qf1, int qf2)
{
    ref_heap_zeroing(1);
    if (g_ready) {
        TurboCodingRelease();
        g_ready = false;
    }
    MODULATION = 1;
    source_length = K;
    length_after_code = 3 * K + 12;
    SYMBOL_NUM = length_after_code;
    f1 = qf1;
    f2 = qf2;
    TurboCodingInit();
    g_ready = true;
}

void ref_release(void)
{
    if (g_ready) TurboCodingRelease();
    g_ready = false;
}

void ref_get_qpp(int *pi) { std::memcpy(pi, index_randomintlvr, sizeof(int) * source_length); }

void ref_get_trellis(int *nextout, int *nextstat, int *lastout, int *laststat)
{
    std::memcpy(nextout, turbo_trellis.mx_nextout, sizeof(int) * 32);
    std::memcpy(nextstat, turbo_trellis.mx_nextstat, sizeof(int) * 16);
    std::memcpy(lastout, turbo_trellis.mx_lastout, sizeof(int) * 32);
    std::memcpy(laststat, turbo_trellis.mx_laststat, sizeof(int) * 16);
}

void ref_encode(int *bits, int *coded) { TurboEnCoding(bits, coded, source_length); }

double ref_max_star(double x, double y) { return E_algorithm(x, y); }

/* BPSK map + AWGN + soft demap exactly as main.cpp:197-202 (uses rand(): seed it here). */
void ref_channel(int *coded, double sigma, unsigned seed, double *llr)
{
    int n = length_after_code;
    std::vector<double> si(n), sq(n), ri(n), rq(n);
    srand(seed);
    module(coded, si.data(), sq.data(), n, 1);
    AWGN(si.data(), ri.data(), sigma, n);
    AWGN(sq.data(), rq.data(), sigma, n);
    demodule(ri.data(), rq.data(), n, llr, 1 / (2 * sigma * sigma), 1);
}

/* The reference's mapper and soft demapper, ITTC/modanddem.cpp:175,674 (M = modu_index). */
void ref_module(int *bits, double *si, double *sq, int n_bits, int M) { module(bits, si, sq, n_bits, M); }
void ref_demodule(double *si, double *sq, int n_sym, double *out, double kf, int M) { demodule(si, sq, n_sym, out, kf, M); }

/* The reference's own TurboDecoding(): N_ITERATION = 15 iterations, mutates llr (x0.5). */
void ref_turbo_decoding(double *llr, int *flow_decoded)
{
    scrub_small_blocks(sizeof(double) * (source_length + M_num_reg + 1));
    TurboDecoding(llr, flow_decoded, 3 * source_length + 4 * M_num_reg);
}

void ref_siso(double *recs, double *La, int terminated, double *LLR, int T)
{
    scrub_small_blocks(sizeof(double) * (T + 1));
    Log_MAP_decoder(recs, La, terminated, LLR, T);
}

/* The loop of TurboDecoding() (log_map.cpp:1202-1265) re-stated around the reference's own
 * Log_MAP_decoder / demultiplex / (de)interleavers / decision so that the iteration count
 * is a parameter and the LLRs (local to TurboDecoding and freed, :1277) can be observed.
 * llr is not mutated. */
void ref_decode_iters(const double *llr, int n_iter, int *bits_out, double *llr1_out,
                      double *llr2_out, double *le_out)
{
    int K = source_length, T = K + M_num_reg, n = 3 * K + 4 * M_num_reg;
    std::vector<double> h(n), yk(4 * T), La(T, 0.0), Le(T, 0.0), LLR(T, 0.0);
    std::vector<int> tmp(T);
    for (int i = 0; i < n; i++) h[i] = llr[i] * 0.5;
    scrub_small_blocks(sizeof(double) * (T + 1));
    demultiplex(h.data(), K, yk.data());
    for (int it = 0; it < n_iter; it++) {
        random_deinterlvr_double(La.data(), Le.data(), index_randomintlvr, K);
        for (int i = K; i < T; i++) La[i] = 0;
        Log_MAP_decoder(yk.data(), La.data(), 1, LLR.data(), T);
        if (llr1_out && it == n_iter - 1) std::memcpy(llr1_out, LLR.data(), sizeof(double) * T);
        for (int i = 0; i < T; i++) Le[i] = LLR[i] - La[i] - 2 * yk[2 * i];
        randominterleaver_double(Le.data(), La.data(), index_randomintlvr, K);
        for (int i = K; i < T; i++) La[i] = 0;
        Log_MAP_decoder(yk.data() + 2 * T, La.data(), 1, LLR.data(), T);
        for (int i = 0; i < T; i++) Le[i] = LLR[i] - La[i] - 2 * yk[2 * T + 2 * i];
        if (bits_out) {
            decision(LLR.data(), T, tmp.data());
            random_deinterlvr_int(bits_out + (size_t)K * it, tmp.data(), index_randomintlvr, K);
        }
    }
    if (llr2_out) std::memcpy(llr2_out, LLR.data(), sizeof(double) * T);
    if (le_out) std::memcpy(le_out, Le.data(), sizeof(double) * T);
}

/* CPU baseline: one codeword per thread (the decode functions only read the globals after
 * init).  Returns wall seconds around the decode calls only. */
double ref_decode_batch(const double *llrs, int n_cb, int n_iter, int *bits_last, int n_threads)
{
    int K = source_length, n = 3 * K + 4 * M_num_reg;
    std::mutex mu;
    int next = 0;
    auto worker = [&]() {
        std::vector<int> bits((size_t)K * n_iter);
        for (;;) {
            int c;
            {
                std::lock_guard<std::mutex> lk(mu);
                c = next++;
            }
            if (c >= n_cb) break;
            ref_decode_iters(llrs + (size_t)c * n, n_iter, bits.data(), 0, 0, 0);
            if (bits_last)
                std::memcpy(bits_last + (size_t)c * K, bits.data() + (size_t)K * (n_iter - 1),
                            sizeof(int) * K);
        }
    };
    if (n_threads < 1) n_threads = 1;
    /* timing leg: run the allocator natively (no zero-fill, ~7 % faster); each worker thread gets
     * its own fresh glibc arena, so the uninitialised tempmax[] reads zero pages anyway */
    ref_heap_zeroing(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; i++) th.emplace_back(worker);
    for (auto &t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    ref_heap_zeroing(1);
    return std::chrono::duration<double>(t1 - t0).count();
}

} /* extern "C" */
