/*
 * turbo_oracle.h -- CPU oracle for the turbo-decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (xinxu27/turbo_decoder_cuda, ITTC/log_map.cpp) used as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Nothing in the product path (turbo_decoder_cuda_b200/, include/) may link or
 * call it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement
 * against the reference's own code compiled in place (oracle/_ref, built by
 * oracle/Makefile from /root/reference/ITTC/{log_map,modanddem}.cpp) and against
 * the golden vectors under tests/golden/ generated from that build
 * (tests/golden/make_golden.py).
 */
#ifndef TURBO_ORACLE_H
#define TURBO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { TDO_ALGO_LOGMAP_LUT = 1, TDO_ALGO_MAXLOG = 2 };

/* (13,15)_8 trellis tables, same layout as TURBO_TRELLIS (ITTC/log_map.h:61-68):
 * nextout[8*4] = {in0,par0,in1,par1} as +-1, nextstat[8*2], lastout[8*4], laststat[8*2]. */
void tdo_gen_trellis(int *nextout, int *nextstat, int *lastout, int *laststat);

/* QPP interleaver pi(i) = (f1*i + f2*i^2) mod K   (ITTC/log_map.cpp:616-624). */
void tdo_qpp_index(int K, int f1, int f2, int *pi);

/* LTE (K,f1,f2) table lookup (TS 36.212 Table 5.1.3-3).
 * Returns 0 on success, -1 if K is not an LTE block size. */
int tdo_lte_qpp_params(int K, int *f1, int *f2);
int tdo_lte_num_sizes(void);
int tdo_lte_size_at(int idx);

/* PCCC encoder, reference mux order: out[3i]=sys, [3i+1]=par1, [3i+2]=par2, then
 * (x,z)x3 tail of RSC1 and (x',z')x3 tail of RSC2  (ITTC/log_map.cpp:451-583). */
void tdo_turbo_encode(const int *bits, int K, const int *pi, int *coded /*3K+12*/);

/* BPSK (+1 for bit 1) over AWGN, LLR = 2 r / sigma^2  (ITTC/main.cpp:174,197-202,
 * ITTC/modanddem.cpp:189-224).  Noise is Box-Muller on a counter-based generator
 * (seeded, reproducible) -- NOT the reference's rand()/CLT noise, which is not
 * reproducible (ITTC/main.cpp:170). */
void tdo_channel_llr(const int *coded, int n, double sigma, unsigned long long seed,
                     unsigned long long stream, double *llr);
double tdo_sigma_from_ebn0(double ebn0_db, int K);

/* max*(x,y) with the reference's 16-entry LUT  (ITTC/log_map.cpp:14-18,779-801). */
double tdo_max_star_lut(double x, double y);

/* One BCJR pass  (ITTC/log_map.cpp:898-1047).  recs = 2*T interleaved (xs,xp)
 * half-LLRs, La[T], LLR[T] out.  algo selects LUT max* or plain max.
 * tempmax_floor: the reference compares against an uninitialised malloc'd
 * tempmax[] (log_map.cpp:925,989); NAN here means "max over states only",
 * a finite value v means tempmax = max(v, max_j alpha).  */
void tdo_siso(const double *recs, const double *La, int terminated, double *LLR,
              int T, int algo, double tempmax_floor);

/* Iterative PCCC decode  (ITTC/log_map.cpp:1146-1280).  Input is NOT mutated
 * (the reference halves it in place, :1202-1205).
 *   llr_in[3K+12]                      channel LLRs, reference mux order
 *   bits_out[n_iter*K]                 hard decisions after every iteration (may be NULL)
 *   llr1_out[T], llr2_out[T]           a-posteriori LLRs of SISO1 (natural order) and
 *                                      SISO2 (interleaved order) of the LAST iteration (may be NULL)
 *   le_out[T]                          extrinsic of SISO2 of the last iteration (may be NULL)  */
void tdo_turbo_decode(const double *llr_in, int K, const int *pi, int n_iter, int algo,
                      int *bits_out, double *llr1_out, double *llr2_out, double *le_out);

/* Same, one codeblock per thread over n_threads host threads.  llr_in is
 * [n_cb][3K+12]; bits_last[n_cb][K] gets the last iteration's decisions.
 * Returns wall seconds spent inside the decode calls (steady clock). */
double tdo_turbo_decode_batch(const double *llr_in, int n_cb, int K, const int *pi,
                              int n_iter, int algo, int *bits_last, int n_threads);

/* ------------------------------------------------------------------------
 * Fixed-point windowed max-log-MAP model.  Bit-exact integer mirror of the
 * s16x2 CUDA kernel (turbo_decoder_cuda_b200/csrc/tdb200_maxlog.cu): same
 * quantisation, same sub-block schedule, same boundary initialisation, same
 * extrinsic scaling -- plain int32 arithmetic with range checks.
 * ---------------------------------------------------------------------- */
typedef struct tdo_fx_params {
    int K;
    int n_iter;
    int sub_len;      /* L: trellis steps per sub-block, multiple of 8, divides K */
    int warmup;       /* guard steps re-run from the neighbouring sub-block (0 = NII only) */
    int frac_bits;    /* LLR quantisation: q = rint(llr * 2^frac_bits) */
    int llr_clip;     /* |q| clamp for channel values */
    int ext_clip;     /* |Le| clamp */
    int ext_scale_q2; /* extrinsic scale in quarters: 3 = 0.75, 4 = 1.0 */
    int early_term;   /* 1: stop when an iteration changes no hard decision and every |a-posteriori| >= et_threshold */
    int et_threshold; /* fixed-point units; values < 1 are treated as 1 */
    int crc_poly;     /* early_term == 2: stop when the natural-order decisions of SISO-1 (second iteration
                         on) divide by x^24 + crc_poly; the delivered bits are then those decisions */
    int logmap;       /* 1: TDB200_ALGO_LOGMAP_S16 -- max* with the linear correction (see turbo_oracle_fx.c) */
    int lm_t4;        /* correction at d = 0 in fixed-point units; 0 = 5 << (frac_bits - 3) */
    int lm_upper;     /* exploration: upper levels of the a-posteriori trees 0 linear / 1 trapezoid / 2 uncorrected */
    int lm_tt, lm_tc; /* exploration: trapezoid parameters */
    int lm_warm_maxlog; /* exploration: warm-up recursions without the correction */
    int lm_upper_off; /* exploration */
    int lm_t4_lam;    /* exploration: T4 of the first a-posteriori level (0 = lm_t4) */
    int lm_exact;     /* exploration: 1 = linear correction on the exact difference, 2 = exact correction (rounded) */
} tdo_fx_params;

/* Returns the number of iterations run.  bits_out[K] final decisions; le_out
 * (may be NULL) gets the last SISO2 extrinsic in natural order (int).  overflow
 * (may be NULL) is set to 1 if any state metric left the int16 range. */
int tdo_fx_decode(const float *llr_in, const int *pi, const tdo_fx_params *p,
                  int *bits_out, int *le_out, int *overflow);

/* ------------------------------------------------------------------------
 * fp32 specification of the sub-block-parallel Log-MAP / max-log kernels
 * (TDB200_ALGO_LOGMAP_F32 / TDB200_ALGO_MAXLOG_F32), turbo_oracle_f32.c.
 * ---------------------------------------------------------------------- */
typedef struct tdo_f32_params {
    int K;
    int n_iter;
    int sub_len;        /* L */
    int warmup;         /* G */
    int logmap;         /* 0: max, 1: max* with the exact correction, 2: max* with the linear correction */
    int early_term;
    float ext_scale;    /* 1.0 (Log-MAP) / 0.75 (max-log) */
    float ext_clamp;    /* |Le| clamp */
    float et_threshold; /* LLR units */
} tdo_f32_params;

/* Returns the iterations run.  bits_out[K]; llr_out / le_out (may be NULL): last SISO-2
 * a-posteriori / scaled extrinsic, K values in SISO-2 (interleaved) order. */
int tdo_f32_decode(const float *llr_in, const int *pi, const tdo_f32_params *p,
                   int *bits_out, float *llr_out, float *le_out);

/* ------------------------------------------------------------------------
 * Mapper / soft demapper (turbo_oracle_mod.c; ITTC/modanddem.cpp).  M = bits per symbol =
 * the reference's modu_index: 1 BPSK, 2 QPSK, 3 8PSK, 4 16QAM, 6 64QAM. */
int tdo_mod_point(int M, int j, double *pi, double *pq);
int tdo_modulate(const int *bits, int n_bits, int M, double *si, double *sq);                 /* module(), :175 */
int tdo_demap_f64(const double *si, const double *sq, int n_sym, int M, double kf, double *out); /* demodule(), :674 */
/* fp32 model of the device demapper and of the 8-bit hand-over to the s16 decoder */
int tdo_demap_f32(const float *si, const float *sq, int n_sym, int M, float kf, float *out);
void tdo_quant_s8(const float *llr, int n, int frac_bits, int clip, signed char *out);

/* ------------------------------------------------------------------------
 * TS 36.212 rate matching for turbo-coded blocks (turbo_oracle_rm.c; parity UNPINNED -- the
 * reference only declares rate_match()/de_rate_match(), ITTC/main.h:23-24).  Turbo-code side in the
 * reference's multiplex order; Ncb <= 0 means the full circular buffer K_w. */
int tdo_rm_geometry(int K, int *R, int *Kpi, int *ND);          /* returns K_w */
void tdo_rm_circular_buffer(int K, int *w /*K_w: multiplex position or -1 for <NULL>*/);
int tdo_rm_k0(int K, int rv, int Ncb);
int tdo_rm_selection(int K, int E, int rv, int Ncb, int *sel /*E*/);
int tdo_rate_match(const int *coded, int K, int E, int rv, int Ncb, int *e_bits);
int tdo_rate_dematch(const double *e_llr, int K, int E, int rv, int Ncb, int accumulate, double *llr);
int tdo_rate_dematch_f32(const float *e_llr, int K, int E, int rv, int Ncb, int accumulate, float *llr);
/* the same with F filler bits at the head of the code block (<NULL> in d0 and d1, 36.212 5.1.3.2.1): never transmitted;
 * the soft inverse sets them (and their parity-1 positions) to `fill` */
void tdo_rm_circular_buffer_f(int K, int F, int *w);
int tdo_rm_selection_f(int K, int E, int rv, int Ncb, int F, int *sel);
int tdo_rate_match_f(const int *coded, int K, int E, int rv, int Ncb, int F, int *e_bits);
int tdo_rate_dematch_f(const double *e_llr, int K, int E, int rv, int Ncb, int F, int accumulate, double fill, double *llr);

/* ------------------------------------------------------------------------
 * LTE CRC24A / CRC24B and code-block segmentation (turbo_oracle_crc.c; TS 36.212 5.1.1, 5.1.2). */
#define TDO_CRC24A_POLY 0x864CFBu
#define TDO_CRC24B_POLY 0x800063u
unsigned tdo_crc24(const unsigned char *bits, int n, unsigned poly);
int tdo_segmentation(int B, int *out /*C, K_plus, K_minus, C_plus, C_minus, F, L*/);

#ifdef __cplusplus
}
#endif
#endif
