// tdb200_internal.h -- shared declarations between the C-ABI translation unit and the kernel
// translation units.  Nothing here is exported.
#pragma once

#include <cuda_runtime.h>

#include <vector>
#include <stdint.h>

#include "tdb200.h"

namespace tdb200 {

constexpr int kStates = 8;  // (13,15)_8 RSC, ITTC/log_map.cpp:28-29
constexpr int kTail = 3;    // M_num_reg

// (13,15)_8 trellis, derived once on the host by rsc_step() in tdb200_api.cu (cf. gen_trellis,
// ITTC/log_map.cpp:281-337):
//   next state for input 0 / 1:      ns0 = {0,4,5,1,2,6,7,3}   ns1 = {4,0,1,5,6,2,3,7}
//   parity bit for input 0 from s:   c0  = {0,0,1,1,1,1,0,0}   (input 1: 1 - c0)
//   previous state reaching s with input 0 / 1: ls0 = {0,3,4,7,1,2,5,6}  ls1 = {1,2,5,6,0,3,4,7}
struct Trellis {
    int ns0[kStates], ns1[kStates], ls0[kStates], ls1[kStates];
    int par0[kStates];  // parity bit (0/1) leaving state s with input 0
};
const Trellis &host_trellis();

// 3-bit packed trellis tables, state 0 in the low bits (checked against the host-derived
// trellis in tdb200_create()).
__host__ __device__ constexpr unsigned pack8(int a, int b, int c, int d, int e, int f, int g, int h)
{
    return a | (b << 3) | (c << 6) | (d << 9) | (e << 12) | (f << 15) | (g << 18) | (h << 21);
}
constexpr unsigned kNs0 = pack8(0, 4, 5, 1, 2, 6, 7, 3);
constexpr unsigned kNs1 = pack8(4, 0, 1, 5, 6, 2, 3, 7);
constexpr unsigned kLs0 = pack8(0, 3, 4, 7, 1, 2, 5, 6);
constexpr unsigned kLs1 = pack8(1, 2, 5, 6, 0, 3, 4, 7);
constexpr unsigned kPar0 = 0x3C;  // bit s = parity leaving state s with input 0: {0,0,1,1,1,1,0,0}

__host__ __device__ constexpr int tb(unsigned t, int s) { return (t >> (3 * s)) & 7; }
// parity as +-1 for input 0 / input 1 from state s (mx_nextout[s*4+1], [s*4+3])
__host__ __device__ constexpr double o0(int s) { return ((kPar0 >> s) & 1) ? 1.0 : -1.0; }
__host__ __device__ constexpr double o1(int s) { return -o0(s); }

// ------------------------------------------------------------------ fp64 reference-order path
constexpr int kRef64Window = 32;  // trellis steps between the alpha vectors the forward sweep keeps
struct Ref64Workspace {
    // per codeblock, T = K+3 doubles each unless noted
    double *xs1, *xp1, *xs2, *xp2;  // half-LLRs after demultiplex (yk_turbo, log_map.cpp:1083-1127)
    double *La, *Le;
    double *ck;           // [n_win][8]: alpha at the start of every window (alpha and beta themselves stay on chip)
    int n_win;            // ceil(T / kRef64Window)
    int max_batch;
};

struct Ref64Args {
    const void *llr;   // [n_cb][3K+12] device, llr_type
    int llr_type;
    int n_cb, K, n_iter;
    const int *pi;     // [K] device
    const int *pi_inv; // [K] device
    Ref64Workspace ws;
    // outputs (device, nullable)
    uint8_t *bits;
    int32_t *bits_iters;
    double *llr1, *llr2, *ext2;
};
cudaError_t launch_ref64_decode(const Ref64Args &a, cudaStream_t st, int *n_launches);

struct Ref64SisoArgs {
    const double *recs, *La;
    double *LLR;
    int terminated, n_cb, T;
    Ref64Workspace ws;
};
cudaError_t launch_ref64_siso(const Ref64SisoArgs &a, cudaStream_t st, int *n_launches);

// ------------------------------------------------------------------ sub-block-parallel fast path
// Geometry of one codeblock: K = P * L, L = 8 * NW.  Sub-block t in [0,P) owns trellis steps
// [tL, (t+1)L); element n = tL + j lives at shared-memory word  j*P + t  ("step-major"), which
// makes both the natural walk (consecutive t -> consecutive banks) and the QPP walk
// (pi(tL+j) = r_j + L*((q_j + A_t + j*B_t) mod P): same row r_j for all t, column a permutation
// polynomial in t) bank-conflict free when 32 | P.
constexpr int kFxNeg = -14000;  // metric of an impossible state (range analysis in DESIGN.md)
// Internal channel-value sources of the packed decoders beyond tdb200_llr_type: received float symbols, demapped inside the
// load stage (tdb200_decode_symbols_batch with BPSK / QPSK): llr = the I plane, sym_q = the Q plane, kf = 1 / (2 sigma^2)
constexpr int kLlrSymBpskF32 = 8, kLlrSymQpskF32 = 9;

struct FastGeom {
    int K, L, P, NW, G;
    int PP;          // row pitch of the step-major arrays in words (P | 1)
    int threads;     // CTA size: P rounded up to a warp multiple
    int n_ckpt;      // alpha checkpoints kept in shared memory per thread: max(NW-2, 0)
    int smem_bytes;
    int resident_ctas;  // CTAs of this geometry the whole device holds at once
    int NP;             // codeblock pairs one CTA decodes side by side (planner rule in tdb200_create; 1 for compile-time geometry)
    int pair_bytes;     // shared memory of one pair's region
};

struct FastArgs {
    const void *llr;  // [n_cb][3K+12] device
    const void *sym_q;  // kLlrSym*: the Q plane
    float kf;           // kLlrSym*: 1 / (2 sigma^2)
    int llr_type;
    int n_cb;
    FastGeom g;
    int n_iter;
    int frac_bits, llr_clip, ext_lim /* Ce+1, multiple of 4 */, q2;
    int early_term, et_threshold;  // early_term: 0 off, 1 decisions + magnitude, 2 CRC of the SISO-1 decisions
    uint32_t crc_poly;             // low 24 bits of the generator (early_term == 2)
    const uint32_t *crc_tab;       // [256] device: byte-wise table of that generator
    const uint32_t *crc_shift;     // [P] device: x^((P-1-t)L) mod g, the weight of sub-block t's remainder
    int logmap;          // 1: TDB200_ALGO_LOGMAP_S16 (max* with the linear correction), 0: max-log
    int lm_t4;           // Log-MAP: correction at |d| = 0 in fixed-point units (5 << (frac_bits - 3))
    uint32_t opaque[4];  // {0xffffffff, 4, 65536, 0xC0000000}: see PassCfg in tdb200_fast_kernel.cuh
    const uint16_t *tab2;  // [L*PP] device: smem word of element pi(tL+j), stored at index j*PP+t
    int prefetch_stride;   // CTAs resident on the device at once (0 = no L2 prefetch of the next pair)
    int pairs_per_cta;     // filled in by launch_fast_s16
    int sm_count;
    // outputs (device, nullable)
    uint8_t *bits;
    int32_t *bits_iters;  // [n_cb][n_iter][K] decisions after every iteration (diagnostic; forces a decision pass per iteration)
    int32_t *iters_used;
    float *llr2, *ext2;  // [n_cb][K+3]
};
// ------------------------------------------------------------------ fp32 sub-block-parallel path
struct F32Args {
    const void *llr;  // [n_cb][3K+12] device
    int llr_type;
    int n_cb;
    FastGeom g;       // same geometry rules as the int16 kernel, one codeblock per CTA
    int n_iter;
    int logmap;       // 0: max, 1: max* with the exact correction, 2: max* with the linear correction
    float ext_scale;  // 1.0 (Log-MAP) or 0.75 (max-log)
    float ext_clamp;  // |Le| clamp (LLR units)
    int early_term;
    float et_threshold;    // LLR units
    const uint16_t *tab2;  // [L*PP] device: smem word of element pi(tL+j), stored at index j*PP+t
    uint8_t *bits;
    int32_t *iters_used;
    float *llr2, *ext2;  // [n_cb][K+3]
};
cudaError_t f32_configure(const FastGeom &g);
cudaError_t launch_f32(const F32Args &a, cudaStream_t st, int *n_launches);
int f32_smem_bytes(const FastGeom &g);

// Compile-time geometry exists for P sub-blocks of 8*NW steps with guard 16, P in {32, 64, 128}, NW in {4, 5, 6}
// (K = 1024 ... 6144 in nine sizes -- the BASELINE size is P=128, NW=6), plus two alternative plans
// for K = 6144: guard 8 (-0.04 dB, +4 %), and 192 sub-blocks of 32 steps.  Everything else runs the generic kernel.
inline bool fast_spec_pn(const FastGeom &g) { return (g.P == 32 || g.P == 64 || g.P == 128) && g.NW >= 4 && g.NW <= 6 && g.G == 16 && g.PP == (g.P | 1); }
inline bool fast_spec128g8(const FastGeom &g) { return g.P == 128 && g.NW == 6 && g.G == 8 && g.PP == 129; }
inline bool fast_spec192(const FastGeom &g) { return g.P == 192 && g.NW == 4 && g.G == 16 && g.PP == 193; }
// every other plan with 2..128 sub-blocks of 32..64 steps and guard 16: P at run time, NW and G compile-time
// (CTAs of up to 128 threads; below 43 sub-blocks several codeblock pairs share one)
inline bool fast_spec_rt(const FastGeom &g) { return !fast_spec_pn(g) && g.P >= 2 && g.P <= 128 && g.NW >= 4 && g.NW <= 8 && g.G == 16 && g.PP == (g.P | 1); }
// 129..192 sub-blocks of 32 or 40 steps: six warps per CTA at 168 registers, two CTAs per SM
inline bool fast_spec_rt192(const FastGeom &g) { return !fast_spec192(g) && g.P > 128 && g.P <= 192 && (g.NW == 4 || g.NW == 5) && g.G == 16 && g.PP == (g.P | 1); }

// Log-MAP kernels with compile-time geometry: 128 sub-blocks of 32 / 40 / 48 steps, guard 16, 24 or 32
inline bool fast_spec_lm(const FastGeom &g) { return g.P == 128 && g.NW >= 4 && g.NW <= 6 && (g.G == 16 || g.G == 24 || g.G == 32) && g.PP == 129; }

// ... and with the sub-block count at run time: 2..128 sub-blocks of 32..64 steps, guard 24 (the auto plan's)
inline bool fast_spec_lm_rt(const FastGeom &g) { return !fast_spec_lm(g) && g.P >= 2 && g.P <= 128 && g.NW >= 4 && g.NW <= 8 && g.G == 24 && g.PP == (g.P | 1); }

cudaError_t fast_s16_configure(FastGeom &g, int sm_count, bool logmap);  // opt in to the dynamic shared memory size
cudaError_t launch_fast_s16(const FastArgs &a, cudaStream_t st, int *n_launches);
int fast_s16_smem_bytes(const FastGeom &g);
int fast_s16_pair_bytes(const FastGeom &g);
bool fast_s16_specialised(const FastGeom &g, bool logmap = false);  // compile-time geometry: one pair per CTA

// ------------------------------------------------------------------ caller side: encoder + channel
struct EncodeArgs {
    const uint8_t *bits;  // [n_cb][K] device, one byte per bit
    uint8_t *coded;       // [n_cb][3K+12] device
    const int *pi;        // [K] device
    int K, n_cb;
    int stride_bits, stride_out, stride_pi;  // shared-memory chunk strides, filled in by launch_encode
    unsigned magic_c, magic_3c;              // ceil(2^32 / C), ceil(2^32 / 3C), C = ceil(K/32)
};
struct ChannelArgs {
    const uint8_t *coded;  // [n] device
    size_t n;
    float sigma;
    unsigned long long seed;
};
cudaError_t launch_encode(const EncodeArgs &a, cudaStream_t st);
cudaError_t launch_channel(const ChannelArgs &a, void *llr, int llr_type, cudaStream_t st);

// ------------------------------------------------------------------ caller side: mapper / soft demapper (tdb200_modem.cu)
struct DemapArgs {
    const void *sym_i, *sym_q;  // [n_llr / modulation] each, device, element type sym_type
    int sym_type;               // TDB200_LLR_F32 / F64 / F16
    void *llr;                  // [n_llr] device, element type llr_type
    int llr_type;               // F64: reference-order fp64 demapper; F32 / F16 / S8: fp32 demapper
    size_t n_llr;               // multiple of 12
    int modulation;             // bits per symbol: 1, 2, 3, 4, 6
    double kf;                  // 1 / (2 sigma^2), ITTC/main.cpp:202
    int frac_bits, clip;        // S8 output: clamp(rint(LLR * 2^frac_bits), +-clip)
};
bool modulation_ok(int M);
cudaError_t launch_modulate(const uint8_t *coded, void *si, void *sq, int sym_type, size_t n_bits, int M, cudaStream_t st);
cudaError_t launch_awgn(const void *x, void *y, int type, size_t n, double sigma, unsigned long long seed, cudaStream_t st);
cudaError_t launch_demap(const DemapArgs &a, cudaStream_t st);

// ------------------------------------------------------------------ caller side: TS 36.212 rate matching (tdb200_ratematch.cu)
struct RmArgs {
    const void *e_llr;  // [n_cb][E] device, element type in_type
    void *llr;          // [n_cb][3K+12] device, element type out_type (== in_type, or S8 for the s16 decoder)
    int in_type, out_type;
    const int *inv;     // [3K+12] device: first transmission index of each multiplex position, -1 if never sent
    int nnn;            // transmitted positions per wrap of the circular buffer
    int NL, E, n_cb, accumulate;
    int frac_bits, clip;
};
// inv[] marker of a filler position (36.212 5.1.3.2.1: d0 / d1 of the F filler bits are <NULL>, never transmitted, known to
// be 0) and the soft value the inverse writes there: a confident 0 (positive = 1 in this library), small enough to survive
// HARQ accumulation in binary16
constexpr int kRmFiller = -2;
constexpr int kRmFillerLlr = -100;
bool build_rm_table(int K, int rv, int Ncb, int F, std::vector<int> &perm, std::vector<int> &inv);
cudaError_t launch_rate_match(const uint8_t *coded, uint8_t *e_bits, const int *perm, int nnn, int NL, int E, int n_cb, cudaStream_t st);
cudaError_t launch_rate_dematch(const RmArgs &a, cudaStream_t st);

// ------------------------------------------------------------------ transport-block side: CRC24A / CRC24B (tdb200_crc.cu)
struct CrcArgs {
    uint8_t *bits;       // [n_cb][K] device, one byte per bit (attach: the last 24 are written)
    int K, n_cb;
    unsigned poly;       // low 24 bits of the generator
    int attach;          // 1: write the parity of the first K-24 bits; 0: divide all K bits
    uint8_t *ok;         // [n_cb] check: remainder == 0 (may be NULL)
    int32_t *remainder;  // [n_cb] check: the remainder (may be NULL)
    int chunk;           // filled in by launch_crc24
    unsigned xpow[5];
};
cudaError_t launch_crc24(const CrcArgs &a, cudaStream_t st);

}  // namespace tdb200
