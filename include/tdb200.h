/*
 * tdb200.h -- C ABI of the B200-native batched LTE turbo decoder.
 *
 * This is the drop-in boundary for ONE path of xinxu27/turbo_decoder_cuda: the
 * iterative PCCC decode
 *     void TurboDecoding(double *flow_for_decode, int *flow_decoded, int flow_length)
 *         ITTC/main.h:20, ITTC/log_map.cpp:1146-1280, called from ITTC/main.cpp:221
 * and its component decoder
 *     void Log_MAP_decoder(double *recs, double *La, int terminated, double *LLR, int len_total)
 *         ITTC/log_map.cpp:898-1047
 * The reference has no FFI layer: its boundary is link-level C++ free functions plus
 * caller-defined globals (ITTC/main.h:6-11).  The entry points below are what a binding for
 * that path would bind; turbo_decoder_cuda_b200/compat/ittc_compat.cpp re-exports the
 * reference's own signatures on top of them.
 *
 * Plain pointers and sizes only; no CUDA or torch types.  `stream` is a cudaStream_t passed as
 * void* (NULL = the legacy default stream).  All functions return a tdb200_status; the text of
 * the last failure on the calling thread is available from tdb200_last_error().
 * There is NO CPU fallback: without a CUDA device every entry point fails with
 * TDB200_ERR_NO_DEVICE.
 */
#ifndef TDB200_H
#define TDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDB200_VERSION 6

typedef enum tdb200_status {
    TDB200_OK = 0,
    TDB200_ERR_INVALID_ARG = 1,
    TDB200_ERR_UNSUPPORTED = 2,
    TDB200_ERR_NO_DEVICE = 3,
    TDB200_ERR_CUDA = 4,
    TDB200_ERR_ALLOC = 5
} tdb200_status;

/* Component-decoder arithmetic. */
typedef enum tdb200_algo {
    /* fp64 Log-MAP with the reference's 16-entry max* LUT, unsegmented recursion, the
     * reference's operation order (replaces Log_MAP_decoder, log_map.cpp:898-1047, bit for
     * bit up to its uninitialised tempmax[]).  The LLR-parity mode. */
    TDB200_ALGO_LOGMAP_F64 = 0,
    /* max-log-MAP in packed 16-bit fixed point (two codeblocks per 32-bit lane), sub-block
     * parallel with boundary-state initialisation.  The throughput mode. */
    TDB200_ALGO_MAXLOG_S16 = 1,
    /* fp32 Log-MAP: max*(x,y) = max + ln(1+e^-|x-y|) evaluated exactly (the reference tabulates the
     * same correction in 16 steps), sub-block parallel with boundary-state initialisation -- the
     * windowed Log-MAP.  One codeblock per CTA; extrinsic scale 1.0 as in the reference.  The three fp32 modes keep
     * the channel LLRs in shared memory as IEEE binary16 (metrics and extrinsics are fp32): inputs are rounded to
     * binary16 and clamped to +-65504 on the way in. */
    TDB200_ALGO_LOGMAP_F32 = 2,
    /* fp32 max-log-MAP on the same structure (extrinsic scale 0.75 by default). */
    TDB200_ALGO_MAXLOG_F32 = 3,
    /* fp32 Log-MAP with the linear correction max(0, 0.24904 (2.5068 - |x-y|)) instead of the exact one:
     * no special-function unit on the critical path, about 2.5x the speed of TDB200_ALGO_LOGMAP_F32 at
     * the same FER.  Bit-exact against its plain-C model. */
    TDB200_ALGO_LINLOGMAP_F32 = 4,
    /* Log-MAP in the packed 16-bit arithmetic of TDB200_ALGO_MAXLOG_S16 (two codeblocks per 32-bit lane, same
     * sub-block schedule, one CTA per codeblock pair for all iterations): every max of the alpha, beta and
     * a-posteriori computations is max*(x,y) = max + c(|x-y|) with the linear correction
     * c = max(0, 0.625 - |x-y|/4), the reference's E_algorithm() table (log_map.cpp:779-801, :14-18) fitted by a
     * line, evaluated on quarter-differences shared by the two max* of a trellis butterfly.  Extrinsic scale 1.0
     * as in the reference, 4 fractional bits, guard 24 by default.  The throughput mode that sits on the
     * reference's BER/FER curve; bit-exact against its integer model (oracle/turbo_oracle_fx.c, logmap = 1). */
    TDB200_ALGO_LOGMAP_S16 = 5
} tdb200_algo;

/* Element type of the channel-LLR input. */
typedef enum tdb200_llr_type {
    TDB200_LLR_F64 = 0, /* the reference's type (double *flow_for_decode) */
    TDB200_LLR_F32 = 1,
    TDB200_LLR_S8 = 2,  /* already quantised: q = LLR * 2^frac_bits, saturated to int8 */
    TDB200_LLR_F16 = 3  /* IEEE binary16: half the PCIe / HBM bytes of float32 */
} tdb200_llr_type;

/* Where caller buffers live. */
typedef enum tdb200_mem {
    TDB200_MEM_HOST = 0,  /* pageable or pinned host memory; copies are issued on `stream` */
    TDB200_MEM_DEVICE = 1 /* device memory of the decoder's GPU */
} tdb200_mem;

typedef struct tdb200_config {
    int K;          /* information bits per codeblock (source_length, ITTC/main.h:6) */
    int f1, f2;     /* QPP parameters (ITTC/main.h:9); 0,0 = look K up in TS 36.212 Table 5.1.3-3 */
    int n_iter;     /* full iterations (N_ITERATION, ITTC/log_map.h:30) */
    int algo;       /* tdb200_algo */
    int sub_block;  /* trellis steps per sub-block / window (multiple of 8 dividing K); 0 = auto.
                       Ignored by TDB200_ALGO_LOGMAP_F64 (always unsegmented). */
    int warmup;     /* guard steps recomputed from the neighbouring sub-block before each
                       sub-block boundary; 0 = next-iteration initialisation only */
    int early_term; /* 1 = stop a codeblock when an iteration leaves every hard decision unchanged AND
                       every a-posteriori magnitude is at least et_threshold (min 2 iterations);
                       2 / 3 (the packed 16-bit decoders, TDB200_ALGO_MAXLOG_S16 and TDB200_ALGO_LOGMAP_S16)
                       = stop when the K hard decisions divide by the
                       CRC24B / CRC24A generator (TS 36.212 5.1.1: a code block of a segmented transport
                       block ends in a CRC24B, an unsegmented one in the transport block's CRC24A) --
                       the stop rule the reference leaves as a placeholder (previous/Decoder.cc:1026,
                       1098-1099).  Checked on the natural-order decisions of SISO-1 from the second
                       iteration on; a block without a valid CRC simply runs all n_iter iterations */
    int et_threshold; /* magnitude test of the stopping rule, fixed-point units, a power of two
                         (1 = decisions only).  0 = default: 2^(frac_bits+3), i.e. |LLR| >= 8 */
    int ext_scale_q2; /* extrinsic scaling in quarters for the max-log modes: 3 = 0.75, 4 = 1.0;
                         0 = default (3 for max-log, 4 for Log-MAP) */
    int frac_bits;  /* fixed-point fractional bits of TDB200_ALGO_MAXLOG_S16 / _LOGMAP_S16 and of
                       TDB200_LLR_S8 input; 0 = default (3 for max-log, 4 for Log-MAP) */
    int ext_clip;   /* TDB200_ALGO_MAXLOG_S16: extrinsic values are clamped to [-(ext_clip+1), ext_clip]
                       (fixed-point units); ext_clip+1 must be a multiple of 4.
                       0 = default: 2^(frac_bits+6) - 1, i.e. |Le| < 64 (511 for TDB200_ALGO_LOGMAP_S16) */
    int device;     /* CUDA device ordinal */
    int max_batch;  /* codeblocks the workspace is sized for per launch; larger batches are
                       processed in chunks.  0 = default for the algo */
} tdb200_config;

/* Optional outputs; any pointer may be NULL.  Buffers live in the `mem` space given to the
 * decode call.  T = K + 3. */
typedef struct tdb200_outputs {
    uint8_t *bits;        /* [n_cb][K]  hard decisions after the last iteration run, natural
                             order, one byte per bit (LLR < 0 -> 0 else 1, log_map.cpp:862-879).
                             With early_term = 1 a codeblock delivers the decisions of ITS OWN stopping
                             iteration (iters_used), independent of the codeblocks it shares a CTA with.
                             With the CRC rule (early_term = 2 / 3) the two codeblocks of a 32-bit lane pair
                             leave together: a block's bits are those after max(iters_used) of the pair. */
    int32_t *bits_iters;  /* [n_cb][n_iter][K] decisions after EVERY iteration, as ints: one
                             codeblock's slab is exactly TurboDecoding's flow_decoded
                             (log_map.cpp:1264).  Rows past an early stop repeat the last one.
                             TDB200_ALGO_LOGMAP_F64 and the packed 16-bit decoders (which then take a
                             decision pass every iteration: a diagnostic output, not the throughput path) */
    int32_t *iters_used;  /* [n_cb] iterations actually run */
    /* a-posteriori / extrinsic LLRs of the LAST iteration, in the algo's native float type
     * (double for LOGMAP_F64, float otherwise; MAXLOG_S16 writes float = q / 2^frac_bits):     */
    void *llr_siso1;      /* [n_cb][T] SISO-1 output, natural order   (LLR_all_turbo, :1230) */
    void *llr_siso2;      /* [n_cb][T] SISO-2 output, interleaved order (:1251)              */
    void *ext_siso2;      /* [n_cb][T] SISO-2 extrinsic, interleaved order (Le_turbo, :1258) */
} tdb200_outputs;

typedef struct tdb200_decoder tdb200_decoder;

/* Fill cfg with the defaults for block size K (8 iterations, TDB200_ALGO_MAXLOG_S16). */
int tdb200_default_config(tdb200_config *cfg, int K);

/* TS 36.212 Table 5.1.3-3.  Returns TDB200_ERR_INVALID_ARG if K is not an LTE block size. */
int tdb200_lte_qpp_params(int K, int *f1, int *f2);

/* Replaces TurboCodingInit() / TurboCodingRelease() (log_map.cpp:349-434,1330-1345) for the
 * decode path: builds the QPP tables, uploads constants, allocates the device workspace. */
int tdb200_create(const tdb200_config *cfg, tdb200_decoder **out);
void tdb200_destroy(tdb200_decoder *dec);

/* Replaces TurboDecoding() for a batch.  llr is [n_cb][3K+12] channel LLRs (positive = bit 1)
 * in the reference's multiplex order: [3i]=systematic, [3i+1]=parity 1, [3i+2]=parity 2 for
 * i<K, then (x,z)x3 tail of encoder 1 and (x',z')x3 tail of encoder 2
 * (log_map.cpp:566-578,1103-1123).  The input is never modified (the reference halves it in
 * place, :1202-1205; the compat wrapper reproduces that side effect).
 * Work is enqueued on `stream`; with TDB200_MEM_DEVICE buffers the call is asynchronous, with
 * TDB200_MEM_HOST buffers it returns after the results have landed in the caller's memory. */
int tdb200_decode_batch(tdb200_decoder *dec, const void *llr, int llr_type, int mem, int n_cb,
                        const tdb200_outputs *out, void *stream);

/* Replaces Log_MAP_decoder() for a batch (TDB200_ALGO_LOGMAP_F64 decoders only): one BCJR pass.
 * recs [n_cb][2T] interleaved (xs,xp) half-LLRs, La [n_cb][T], LLR out [n_cb][T], doubles. */
int tdb200_siso_batch(tdb200_decoder *dec, const double *recs, const double *La, int terminated,
                      double *LLR, int mem, int n_cb, void *stream);

/* ---- the caller side of the path, for harnesses that keep everything on the device -------------
 * Replaces TurboEnCoding(int *source, int *coded_source, int source_length) (ITTC/main.h:14,
 * log_map.cpp:700-730) for a batch: bits [n_cb][K] (one byte per bit) -> coded [n_cb][3K+12] in the
 * reference's multiplex order, (13,15)_8 PCCC with trellis termination.  Bit-exact with the reference. */
int tdb200_encode_batch(tdb200_decoder *dec, const uint8_t *bits, uint8_t *coded, int mem, int n_cb, void *stream);

/* Replaces module() (BPSK) + AWGN() + demodule() (ITTC/main.cpp:197-202): coded [n_cb][3K+12] bits ->
 * llr [n_cb][3K+12] = 2 r / sigma^2, r = (2c-1) + sigma * n.  n is standard normal from Philox4x32-10,
 * a pure function of (seed, element index) -- the reference's 12-term CLT noise seeded from time() is
 * not reproducible, so the agreement is statistical.  llr_type: TDB200_LLR_F32 or TDB200_LLR_F64. */
int tdb200_channel_batch(tdb200_decoder *dec, const uint8_t *coded, void *llr, int llr_type, int mem, int n_cb,
                         double sigma, uint64_t seed, void *stream);

/* ---- higher-order mapping: the reference's module()/demodule() pair (ITTC/modanddem.cpp:175,674) ----
 * `modulation` is the reference's modu_index = bits per symbol (MODULATION, ITTC/main.cpp:29):
 * 1 BPSK, 2 QPSK, 3 8PSK, 4 16QAM, 6 64QAM, with the reference's own constellation tables
 * (modanddem.cpp:7-71) and bit order.  Symbols are planar like the reference's symbol_i / symbol_q:
 * two arrays [n_cb][(3K+12)/modulation] of sym_type (TDB200_LLR_F32, _F64 or _F16). */
typedef enum tdb200_modulation {
    TDB200_MOD_BPSK = 1, TDB200_MOD_QPSK = 2, TDB200_MOD_8PSK = 3, TDB200_MOD_16QAM = 4, TDB200_MOD_64QAM = 6
} tdb200_modulation;

/* Replaces module(int *a, double *outi, double *outq, int N, int modu_index) for a batch:
 * coded [n_cb][3K+12] bits (one byte per bit) -> constellation points.  Exact (table look-up). */
int tdb200_modulate_batch(tdb200_decoder *dec, const uint8_t *coded, void *sym_i, void *sym_q, int sym_type,
                          int mem, int n_cb, int modulation, void *stream);

/* Replaces AWGN(double *in, double *out, double sigma, int len) (ITTC/log_map.cpp:1388): y = x + sigma*n
 * over n values of `type`; n is standard normal from Philox4x32-10, a pure function of (seed, index). */
int tdb200_awgn_batch(tdb200_decoder *dec, const void *x, void *y, int type, int mem, size_t n,
                      double sigma, uint64_t seed, void *stream);

/* Replaces demodule(double *symbol_i, double *symbol_q, int symbol_len, double *out, double Kf,
 * int modu_index) for a batch: received symbols -> llr [n_cb][3K+12], max-log,
 * LLR_b = -Kf * (min_{points with bit b = 1} d - min_{points with bit b = 0} d), Kf = 1/(2 sigma^2)
 * (ITTC/main.cpp:202).  llr_type TDB200_LLR_F64: fp64 in the reference's order of operations,
 * bit-identical to demodule().  TDB200_LLR_F32 / _F16: the same metric in fp32 (per axis for the
 * product constellations).  TDB200_LLR_S8: the fp32 metric quantised to the throughput decoder's
 * fixed-point channel values, clamp(rint(LLR * 2^frac_bits), +-127) -- what tdb200_decode_batch
 * computes itself from float LLRs, so decoding these bytes equals decoding the float LLRs. */
int tdb200_demap_batch(tdb200_decoder *dec, const void *sym_i, const void *sym_q, int sym_type,
                       void *llr, int llr_type, int mem, int n_cb, int modulation, double kf, void *stream);

/* The same two stages on flat arrays of any length (n_bits / n_llr a multiple of `modulation`), for
 * rows that are not 3K+12 long -- rate-matched blocks: tdb200_rate_match_batch -> tdb200_modulate_flat
 * on the way out, tdb200_demap_flat -> tdb200_decode_rm_batch on the way in. */
int tdb200_modulate_flat(tdb200_decoder *dec, const uint8_t *bits, void *sym_i, void *sym_q, int sym_type,
                         int mem, size_t n_bits, int modulation, void *stream);
int tdb200_demap_flat(tdb200_decoder *dec, const void *sym_i, const void *sym_q, int sym_type,
                      void *llr, int llr_type, int mem, size_t n_llr, int modulation, double kf, void *stream);

/* demodule() + TurboDecoding() in one call (ITTC/main.cpp:202,221): received symbols in, decisions
 * out.  The demapped values stay on the device in the decoder's own input format (fp64 for
 * TDB200_ALGO_LOGMAP_F64, 8-bit fixed point for TDB200_ALGO_MAXLOG_S16, float for the fp32 modes),
 * so a host caller moves 2*sizeof(sym)/modulation bytes per LLR over PCIe instead of sizeof(llr).
 * Results equal tdb200_demap_batch followed by tdb200_decode_batch.  `out` as for tdb200_decode_batch. */
int tdb200_decode_symbols_batch(tdb200_decoder *dec, const void *sym_i, const void *sym_q, int sym_type,
                                int mem, int n_cb, int modulation, double kf,
                                const tdb200_outputs *out, void *stream);

/* ---- TS 36.212 rate matching: the stage the reference declares and never wrote -------------------
 * void rate_match(int *input, int in_len, int *output, int out_len) and
 * void de_rate_match(double *input, double *output, int in_len, int out_len)  (ITTC/main.h:23-24,
 * call sites commented out at ITTC/main.cpp:196,204).  Implemented per TS 36.212: tail-bit
 * multiplexing into d0/d1/d2 (5.1.3.2.2), sub-block interleavers, circular buffer and bit selection
 * from k0(rv) with <NULL> pruning (5.1.4.1).  The turbo-code side keeps the reference's multiplex
 * order [n_cb][3K+12]; the channel side is [n_cb][E].  rv = redundancy version 0..3; ncb = soft
 * buffer size N_cb (0: the full circular buffer K_w = 3 * 32 * ceil((K+4)/32)).
 * FILLER BITS (5.1.2: the first F bits of a transport block's first code block): 5.1.3.2.1 turns d0_k and d1_k,
 * k < F, into <NULL>s that are never transmitted.  tdb200_set_filler_bits(dec, F) declares them for every code block
 * that goes through this handle afterwards (use one handle for the first blocks, F > 0, and one for the rest): the
 * bit selection then skips those 2F positions like the interleaver's dummy bits, and the soft inverse writes a fixed
 * confident "0" (-100; -127 in the 8-bit format) at the F systematic and F parity-1 positions -- the receiver knows
 * them (the encoder starts in state 0 and stays there while it reads zeros).  F = 0 (the default): no filler bits.
 * Pinned by a hand-derived K = 40 known answer (tests/test_oracle.py). */
int tdb200_set_filler_bits(tdb200_decoder *dec, int F);

int tdb200_rate_match_batch(tdb200_decoder *dec, const uint8_t *coded, uint8_t *e_bits, int mem, int n_cb,
                            int E, int rv, int ncb, void *stream);

/* Soft inverse: e_llr [n_cb][E] -> llr [n_cb][3K+12], both of llr_type.  Bits sent more than once
 * (E beyond one wrap of the buffer) are summed in transmission order, bits not sent come out as 0.
 * accumulate != 0 adds to what `llr` already holds (HARQ combining of a retransmission with another
 * rv).  Sums are fp32 (fp64 for TDB200_LLR_F64, saturating integers for TDB200_LLR_S8). */
int tdb200_rate_dematch_batch(tdb200_decoder *dec, const void *e_llr, void *llr, int llr_type, int mem, int n_cb,
                              int E, int rv, int ncb, int accumulate, void *stream);

/* de_rate_match() + TurboDecoding() in one call (ITTC/main.cpp:204,221): rate-matched LLRs in,
 * decisions out; the de-rate-matched values stay on the device (for TDB200_ALGO_MAXLOG_S16 already
 * in the decoder's 8-bit channel format).  Results equal tdb200_rate_dematch_batch followed by
 * tdb200_decode_batch.  `out` as for tdb200_decode_batch. */
int tdb200_decode_rm_batch(tdb200_decoder *dec, const void *e_llr, int llr_type, int mem, int n_cb,
                           int E, int rv, int ncb, const tdb200_outputs *out, void *stream);

/* ---- transport-block stage above the decoder (TS 36.212 5.1.1 / 5.1.2) ----------------------------
 * The reference has a placeholder only (previous/Decoder.cc:1026 "stoprule ... 1=CRC", :1098-1099).
 * `which`: TDB200_CRC24A (transport block, g = 0x1864CFB) or TDB200_CRC24B (code block, g = 0x1800063).
 * bits are [n_rows][row_bits], one byte per bit (the decoder's own output format); row_bits = 0 means
 * the handle's K (code blocks), any other length >= 25 serves transport blocks. */
typedef enum tdb200_crc { TDB200_CRC24A = 0, TDB200_CRC24B = 1 } tdb200_crc;

/* Overwrites the last 24 bits of every row with the CRC of the bits before them. */
int tdb200_crc24_attach_batch(tdb200_decoder *dec, uint8_t *bits, int row_bits, int which, int mem, int n_rows, void *stream);

/* ok[c] = 1 if row c divides by the generator (payload + CRC intact); remainder (may be NULL) gets the
 * 24-bit remainder.  Runs on decoded blocks without leaving the device: the ACK/NACK of a code block. */
int tdb200_crc24_check_batch(tdb200_decoder *dec, const uint8_t *bits, int row_bits, int which, uint8_t *ok,
                             int32_t *remainder, int mem, int n_rows, void *stream);

/* Code-block segmentation of a transport block of B bits (its CRC24A included), 5.1.2: C blocks,
 * C_plus of size K_plus and C_minus of size K_minus, F filler bits at the head of the first block,
 * L = 24 CRC24B bits per block when C > 1.  Host-only arithmetic (no device needed). */
typedef struct tdb200_seg_info { int C, K_plus, K_minus, C_plus, C_minus, F, L; } tdb200_seg_info;
int tdb200_segmentation(int B, tdb200_seg_info *info);

/* Introspection (what the plan resolved to). */
typedef struct tdb200_plan_info {
    int K, f1, f2, n_iter, algo;
    int sub_block;       /* L */
    int n_sub_blocks;    /* P = K / L */
    int warmup;
    int cb_per_cta;      /* codeblocks one CTA decodes concurrently */
    int threads_per_cta;
    int smem_bytes;      /* dynamic shared memory per CTA */
    int max_batch;
    int sm_count;
    int kernel_launches_last_call; /* kernels launched by the most recent decode call */
} tdb200_plan_info;
int tdb200_get_plan(const tdb200_decoder *dec, tdb200_plan_info *info);

/* Measurement aid (no reference counterpart): issue rate of the add-compare-select instruction mix on `device`, in
 * thread-operations per clock per SM (128 = one warp-instruction per clock on each of the four sub-partitions).
 * mix 0: VIADDMNMX.S16x2 alone (ALU pipe), 1: VIADD.16x2 alone (fma-heavy pipe), 2: both interleaved -- the mix of the
 * recursions.  bench.py turns it into the measured denominator of its ALU roofline. */
int tdb200_ubench_issue_rate(int device, int mix, double *thread_ops_per_clk_per_sm);

const char *tdb200_last_error(void);
const char *tdb200_status_string(int status);

#ifdef __cplusplus
}
#endif
#endif /* TDB200_H */
