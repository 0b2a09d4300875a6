// tdb200_ref64.cu -- fp64 reference-order Log-MAP decoder (TDB200_ALGO_LOGMAP_F64).
//
// The LLR-parity mode: an unsegmented BCJR that performs the reference's floating-point
// operations on the same operands in the same order as
//     TurboDecoding()    ITTC/log_map.cpp:1146-1280
//     Log_MAP_decoder()  ITTC/log_map.cpp:898-1047   (gamma :962-972, alpha :975-1001,
//                                                      beta :1004-1021, LLR :1024-1039)
//     E_algorithm()      ITTC/log_map.cpp:779-801, LUT :14-18
// so its a-posteriori/extrinsic LLRs agree with the CPU code to rounding noise (the only
// non-reproducible term in the reference is its uninitialised tempmax[], :925/:989; like the
// oracle this kernel normalises by max_j alpha_j).
//
// Mapping (nothing like the reference's loops).  The recursion cannot be cut into sub-blocks here
// (that would change results), so one decode is 32 dependent sweeps of K+3 steps and the kernel is
// bound by the LATENCY of one trellis step; everything is arranged around that chain
// (tools/ubench_fp64.cu: DADD 8 clk, compare+select 14, 64-bit shuffle 26, shared round trip 35):
//   * one warp owns four codeblocks for the whole decode, ONE LANE PER TRELLIS STATE;
//   * alpha and beta never leave the SM: the forward sweep keeps one alpha vector per 32-step
//     window (64 B per window in an L2-resident scratch), the backward sweep re-creates a window's
//     alpha vectors (bit-identical: same operations on the same operands) into shared memory, runs
//     beta over the window, and the a-posteriori sums of the window are then folded with one LANE
//     PER TRELLIS POSITION (E_algorithm_seq is a serial fold over the states, :817-829);
//   * an alpha step exchanges the un-normalised metrics ONCE through shared memory: every lane
//     reads all eight (for max_j, :987-993) and its two predecessors, and subtracts the normaliser
//     itself -- one round trip instead of a 3-round shuffle butterfly plus a predecessor exchange;
//   * max* looks its correction up in a 58-entry HASHED table: the exponent and top three mantissa
//     bits of |x-y| select an interval that holds at most one of the reference's 16 breakpoints,
//     one exact comparison against that breakpoint picks the value -- the same result as the
//     reference's linear scan for every double (tests/test_host_logic.py restates and checks it), in 10 instructions
//     instead of a 15-compare tree;
//   * the (xs,xp,La) inputs of the next window arrive by cp.async while the current one runs.
// Compile with -fmad=false: products here are exact (x * +-1, x * 0.5) so contraction would not
// change results, but the flag keeps that a non-question.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "tdb200_internal.h"

namespace tdb200 {
namespace {

constexpr double kInfty = 1E20;  // ITTC/log_map.h:72-74

constexpr int kW = kRef64Window;           // steps per window
constexpr int kRow = 9;                    // doubles per alpha/beta row: 8 states + 1 (lane stride 18 banks: conflict-free position reads)
constexpr int kCbStride = kW * kRow + 8;   // doubles per codeblock region; 592 words = 16 (mod 32): the two codeblocks of a half-warp do not collide
constexpr int kGtStride = kW * 4 + 8;      // branch-metric table of one codeblock, same staggering
constexpr int kLutEntries = 58;            // |d| < 2^-4, 7 binades x 8 sub-intervals, |d| >= 8

struct Smem {  // one warp per CTA
    double raw[2][3][4][kW];     // staged xs, xp, La of two windows (cp.async double buffer)
    double gt[4][kGtStride];     // branch metrics of the window: [step]{(-s-q)-h, (-s+q)-h, (s-q)+h, (s+q)+h}, h = La/2 (gama_Log, :962-972)
    double aw[4][kCbStride];     // alpha_i of the window, row = step, [state]
    double bw[4][kCbStride];     // beta_{i+1} of the window
    double mw[4][kW + 2];        // normaliser max(0, max_j alpha_j) of step i+1 (tempmax[], :987-993); +2: the four codeblocks' words in different banks
    double ex[2][4][8];          // alpha exchange, double-buffered by step parity
    double lut[2][kLutEntries + 2];  // [0]: the breakpoint inside an interval, [1]: the value below it; the value from it on is the next entry's
};

// The correction table of E_algorithm, hashed by the leading bits of d (see the header).  Entry e:
// 0 covers d < 2^-4, 1..56 the intervals 2^(b-4) * [1 + s/8, 1 + (s+1)/8), 57 covers d >= 8.
// No breakpoint is a multiple of 2^(b-7), so "value from the breakpoint on" = "value at the start
// of the next interval": a breakpoint and one value per interval are enough.  They sit in two planes of 8-byte words
// (a 16-byte {breakpoint, value} load measured 4 % slower: the comparison waits for the whole quad).
__device__ void build_lut(double (*lut)[kLutEntries + 2], int lane)
{
    const double idx[16] = {0.0, 0.08824, 0.19587, 0.31026, 0.43275, 0.56508, 0.70963, 0.86972,
                            1.0502, 1.2587, 1.5078, 1.8212, 2.2522, 2.9706, 3.6764, 4.3758};  // :14-16
    const double val[16] = {0.69315, 0.65, 0.6, 0.55, 0.5, 0.45, 0.4, 0.35,
                            0.3, 0.25, 0.2, 0.15, 0.1, 0.05, 0.025, 0.0};  // :17-18; from 4.3758 on the result is 0 (:784-787)
    for (int e = lane; e < kLutEntries + 2; e += 32) {
        double bp = 1e300, below;
        if (e == 0) below = val[0];
        else if (e >= kLutEntries - 1) below = 0.0;
        else {
            const int b = (e - 1) >> 3, s = (e - 1) & 7;
            const double base = 1.0 / (double)(1 << 4) * (double)(1 << b);
            const double lo = base * (1.0 + 0.125 * s), hi = lo + base * 0.125;
            int k = 0;
            for (int t = 1; t < 16; t++) if (idx[t] <= lo) k = t;  // region of lo: [idx[k], idx[k+1])
            below = val[k];
            if (k < 15 && idx[k + 1] < hi) bp = idx[k + 1];
        }
        lut[0][e] = bp; lut[1][e] = below;
    }
}

// max*(x,y), ITTC/log_map.cpp:779-801
__device__ __forceinline__ double max_star(double x, double y, const double (*lut)[kLutEntries + 2])
{
    const double diff = y - x;  // d = (y-x) > 0 ? (y-x) : (x-y) = |diff|
    int e = ((__double2hiint(diff) & 0x7fffffff) >> 17) - ((1023 - 4) * 8 - 1);
    e = min(max(e, 0), kLutEntries - 1);
    const double bp = lut[0][e], below = lut[1][e], above = lut[1][e + 1];  // three 8-byte loads: the breakpoint is back first
    return (x > y ? x : y) + (fabs(diff) < bp ? below : above);
}

__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }
__device__ __forceinline__ double dmax(double a, double b) { return a < b ? b : a; }

__device__ __forceinline__ double load_llr_half(const void *p, int type, size_t idx)
{
    // flow_for_decode[i] *= 0.5, ITTC/log_map.cpp:1202-1205 (done on a copy)
    if (type == TDB200_LLR_F64) return static_cast<const double *>(p)[idx] * 0.5;
    if (type == TDB200_LLR_F32) return static_cast<double>(static_cast<const float *>(p)[idx]) * 0.5;
    if (type == TDB200_LLR_F16) return static_cast<double>(__half2float(static_cast<const __half *>(p)[idx])) * 0.5;
    return static_cast<double>(static_cast<const int8_t *>(p)[idx]) * 0.0625;  // S8, 3 fractional bits
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Inputs of one BCJR pass for the four codeblocks of this warp (planar, stride sT per codeblock).
struct PassIn {
    const double *xs, *xp, *La;
    size_t sT;
    double *ck;   // alpha at the start of every window: [4][n_win][8]
    size_t sCk;   // = 8 * n_win
    int T, terminated;
};

// Window w of the three input arrays -> sm.raw[buf] (lane (g,j): steps j, j+8, j+16, j+24 of codeblock g).
__device__ __forceinline__ void stage_window(Smem &sm, const PassIn &p, int w, int buf, int g, int j)
{
    if (w >= 0) {
        const int i0 = w * kW;
        const double *src[3] = {p.xs + g * p.sT, p.xp + g * p.sT, p.La + g * p.sT};
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int r = 0; r < kW / 8; r++) {
                const int u = j + 8 * r;
                if (i0 + u < p.T) cp_async8(&sm.raw[buf][k][g][u], src[k] + i0 + u);
            }
    }
    cp_async_commit();
}

// The four branch metrics of every step of the staged window (each lane converts what it staged itself).
__device__ __forceinline__ void gamma_window(Smem &sm, int buf, int g, int j)
{
#pragma unroll
    for (int r = 0; r < kW / 8; r++) {
        const int u = j + 8 * r;
        const double s = sm.raw[buf][0][g][u], q = sm.raw[buf][1][g][u], h = sm.raw[buf][2][g][u] * 0.5;
        double2 *o = reinterpret_cast<double2 *>(&sm.gt[g][u * 4]);
        o[0] = make_double2((-s - q) - h, (-s + q) - h);  // input 0, parity -1 / +1, :967-968
        o[1] = make_double2((s - q) + h, (s + q) + h);    // input 1, :969-970
    }
}

// Steps [0, wlen) of the window, forward.  On entry al = alpha_i(j) of the window's first step and
// (a0, a1) = alpha_i of the two predecessors of state j; on exit the same for the step after the
// window.  STORE: keep alpha_i (rows of sm.aw) and the normalisers (sm.mw) for the beta / LLR phase.
// c0 / c1: this lane's columns of a step's branch-metric row (gamma from ls0 with input 0, from ls1 with input 1).
template <bool STORE>
__device__ __forceinline__ void alpha_window(Smem &sm, int g, int j, int wlen, double &al, double &a0, double &a1,
                                             int ls0, int ls1, int c0, int c1)
{
    const double *gt = sm.gt[g];
    double g0 = gt[c0], g1 = gt[c1];
    if (STORE) sm.aw[g][j] = al;
#pragma unroll 2
    for (int u = 0; u < wlen; u++) {
        const double v = max_star(g0 + a0, g1 + a1, sm.lut);
        double *ex = sm.ex[u & 1][g];
        ex[j] = v;
        // the next step's branch metrics, in the shadow of the exchange
        const int un = min(u + 1, kW - 1);
        g0 = gt[un * 4 + c0];
        g1 = gt[un * 4 + c1];
        __syncwarp();
        const double2 *e2 = reinterpret_cast<const double2 *>(ex);
        const double2 v01 = e2[0], v23 = e2[1], v45 = e2[2], v67 = e2[3];
        const double x0 = ex[ls0], x1 = ex[ls1];
        const double m = dmax(dmax(dmax(v01.x, v01.y), dmax(v23.x, v23.y)), dmax(dmax(v45.x, v45.y), dmax(v67.x, v67.y)));
        // tempmax[i+1] = max(0, max_j alpha_j): the reference compares against an uninitialised
        // tempmax[] (:925,:989); on a clean (zero-filled) heap that is this, which is what oracle/ pins
        const bool neg = m < 0.0;
        a0 = neg ? x0 : x0 - m;  // :996-999
        a1 = neg ? x1 : x1 - m;
        al = neg ? v : v - m;
        if (STORE) {
            if (j == 0) sm.mw[g][u] = neg ? 0.0 : m;
            if (u + 1 < wlen) sm.aw[g][(u + 1) * kRow + j] = al;
        }
    }
}

// Steps [0, wlen) of the window, backward.  be = beta_{i+1}(j) of the window's last step on entry,
// beta_i(j) of its first step on exit.  Rows of sm.bw receive beta_{i+1}.
__device__ __forceinline__ void beta_window(Smem &sm, int g, int j, int wlen, double &be, int ns0, int ns1, int c0, int c1)
{
    const double *gt = sm.gt[g], *mw = sm.mw[g];
    sm.bw[g][(wlen - 1) * kRow + j] = be;
    double g0 = gt[(wlen - 1) * 4 + c0], g1 = gt[(wlen - 1) * 4 + c1], m = mw[wlen - 1];
#pragma unroll 2
    for (int u = wlen - 1; u >= 0; u--) {
        const double tx = g0 + shfl8(be, ns0);
        const double ty = g1 + shfl8(be, ns1);
        const double mm = m;
        const int un = max(u - 1, 0);  // the next step's operands, ahead of this step's chain
        g0 = gt[un * 4 + c0];
        g1 = gt[un * 4 + c1];
        m = mw[un];
        be = max_star(tx, ty, sm.lut) - mm;  // :1004-1021
        if (u >= 1) sm.bw[g][(u - 1) * kRow + j] = be;
    }
}

// LLR of the window, :1024-1039: lane = trellis position.  E_algorithm_seq (:817-829) is a serial fold over the
// eight states -- two chains of seven max* per position -- so the four codeblocks of the warp are folded side by
// side (eight independent chains per lane) to fill the latency of one max*.
template <class Emit>
__device__ __forceinline__ void fold_window(Smem &sm, int buf, int lane, int wlen, int i0, Emit &emit)
{
    if (lane >= wlen) return;
    const int tgt = emit.target(i0 + lane);
    double m0[4], m1[4];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
        const int l0 = tb(kLs0, jj), l1 = tb(kLs1, jj);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double *gt = &sm.gt[c][lane * 4], *a = &sm.aw[c][lane * kRow], *b = &sm.bw[c][lane * kRow];
            const double t0 = (gt[o0(l0) > 0 ? 1 : 0] + a[l0]) + b[jj];  // :1028-1030
            const double t1 = (gt[o1(l1) > 0 ? 3 : 2] + a[l1]) + b[jj];  // :1032-1034
            if (jj == 0) { m0[c] = t0; m1[c] = t1; }
            else { m0[c] = max_star(m0[c], t0, sm.lut); m1[c] = max_star(m1[c], t1, sm.lut); }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
        emit(c, i0 + lane, tgt, m1[c] - m0[c], sm.raw[buf][2][c][lane], sm.raw[buf][0][c][lane]);  // :1038
}

// One BCJR pass (Log_MAP_decoder, :898-1047) for the four codeblocks of this warp.
template <class Emit>
__device__ void siso_pass(Smem &sm, const PassIn &p, int lane, Emit &emit)
{
    const int g = lane >> 3, j = lane & 7;
    const int T = p.T, n_win = (T + kW - 1) / kW;
    const int ls0 = tb(kLs0, j), ls1 = tb(kLs1, j), ns0 = tb(kNs0, j), ns1 = tb(kNs1, j);
    // this lane's columns of a branch-metric row: parity signs of the branches ENTERING state j ...
    const int oa0 = o0(ls0) > 0 ? 1 : 0, oa1 = o1(ls1) > 0 ? 3 : 2;
    // ... and of the branches LEAVING it
    const int ob0 = o0(j) > 0 ? 1 : 0, ob1 = o1(j) > 0 ? 3 : 2;
    double *ck = p.ck + g * p.sCk + j;

    // ---- alpha forward, :975-1001: only the window-start vectors are kept
    double al = (j == 0) ? 0.0 : -kInfty;  // :943-948
    double a0 = (ls0 == 0) ? 0.0 : -kInfty, a1 = (ls1 == 0) ? 0.0 : -kInfty;
    stage_window(sm, p, 0, 0, g, j);
    for (int w = 0; w < n_win; w++) {
        ck[(size_t)w * 8] = al;
        if (w == n_win - 1) break;  // the last window is re-created below anyway
        stage_window(sm, p, w + 1, (w + 1) & 1, g, j);
        cp_async_wait<1>();
        gamma_window(sm, w & 1, g, j);
        __syncwarp();
        alpha_window<false>(sm, g, j, kW, al, a0, a1, ls0, ls1, oa0, oa1);
        __syncwarp();
    }
    // window n_win-1 is in flight (or, with a single window, staged by the first call)
    // ---- backward: re-create alpha per window, beta :1004-1021, LLR :1024-1039
    double be = (j == 0) ? 0.0 : (p.terminated ? -kInfty : 0.0);  // :944-959
    double ck_next = 0.0;  // alpha at the start of window w-1, fetched one window ahead
    for (int w = n_win - 1; w >= 0; w--) {
        const int buf = w & 1, wlen = min(kW, T - w * kW);
        stage_window(sm, p, w - 1, (w - 1) & 1, g, j);
        cp_async_wait<1>();
        gamma_window(sm, buf, g, j);
        __syncwarp();
        if (w < n_win - 1) {  // otherwise (al, a0, a1) are what the forward sweep left
            al = ck_next;
            a0 = shfl8(al, ls0);
            a1 = shfl8(al, ls1);
        }
        if (w >= 1) ck_next = ck[(size_t)(w - 1) * 8];
        alpha_window<true>(sm, g, j, wlen, al, a0, a1, ls0, ls1, oa0, oa1);
        __syncwarp();
        beta_window(sm, g, j, wlen, be, ns0, ns1, ob0, ob1);
        __syncwarp();
        fold_window(sm, buf, lane, wlen, w * kW, emit);
        __syncwarp();
    }
    cp_async_wait<0>();
}

struct DecodeEmit {
    const Ref64Args &a;
    double *La_next;  // a-priori array of the next pass (the warp's four codeblocks)
    const int *scat;  // where this pass's position i sits in the next pass's order
    size_t sT;
    int cb0, T, K, it, siso;
    bool last;
    __device__ __forceinline__ int target(int i) const { return i < K ? scat[i] : -1; }
    __device__ __forceinline__ void operator()(int c, int i, int tgt, double L, double la, double xs) const
    {
        // extrinsic, :1234-1238 / :1255-1259, stored where the next pass reads it as a-priori value:
        // SISO-1 -> SISO-2 La[i'] = Le[pi(i')] (randominterleaver_double, :1242), i.e. i' = pi^-1(i);
        // SISO-2 -> SISO-1 La[pi(i)] = Le[i] (random_deinterlvr_double, :1221); tails stay 0 (:1224-1227)
        const double le = L - la - 2 * xs;
        if (tgt >= 0) La_next[c * sT + tgt] = le;
        const int cb = cb0 + c;
        if (cb >= a.n_cb) return;
        if (siso == 1 && tgt >= 0) {  // decision + deinterleave, :1261-1264 (tgt = pi(i))
            const int bit = (L < 0) ? 0 : 1;
            if (a.bits_iters) a.bits_iters[((size_t)cb * a.n_iter + it) * K + tgt] = bit;
            if (last && a.bits) a.bits[(size_t)cb * K + tgt] = (uint8_t)bit;
        }
        if (last) {
            if (siso == 0 && a.llr1) a.llr1[(size_t)cb * T + i] = L;
            if (siso == 1 && a.llr2) a.llr2[(size_t)cb * T + i] = L;
            if (siso == 1 && a.ext2) a.ext2[(size_t)cb * T + i] = le;
        }
    }
};

__global__ void __launch_bounds__(32) ref64_decode_kernel(Ref64Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int lane = threadIdx.x;
    const int cb0 = blockIdx.x * 4;
    if (cb0 >= a.n_cb) return;
    const int K = a.K, T = K + kTail, NL = 3 * K + 4 * kTail;
    const size_t sT = T;
    const Ref64Workspace &w = a.ws;
    double *xs1 = w.xs1 + cb0 * sT, *xp1 = w.xp1 + cb0 * sT, *xs2 = w.xs2 + cb0 * sT, *xp2 = w.xp2 + cb0 * sT;
    double *La = w.La + cb0 * sT, *Ln = w.Le + cb0 * sT;  // a-priori values of this pass / of the next
    build_lut(sm.lut, lane);

    // ---- x0.5 and demultiplex, :1202-1209, :1083-1127
    for (int c = 0; c < 4; c++) {
        const int cb = cb0 + c;
        const bool valid = cb < a.n_cb;
        const size_t base = (size_t)(valid ? cb : cb0) * NL;
        for (int i = lane; i < T; i += 32) {
            double s1, p1, s2, p2;
            if (i < K) {
                s1 = load_llr_half(a.llr, a.llr_type, base + 3 * i);
                p1 = load_llr_half(a.llr, a.llr_type, base + 3 * i + 1);
                p2 = load_llr_half(a.llr, a.llr_type, base + 3 * i + 2);
                s2 = load_llr_half(a.llr, a.llr_type, base + 3 * (size_t)a.pi[i]);
            } else {
                const int m = i - K;
                s1 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * m);
                p1 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * m + 1);
                s2 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * kTail + 2 * m);
                p2 = load_llr_half(a.llr, a.llr_type, base + 3 * K + 2 * kTail + 2 * m + 1);
            }
            xs1[c * sT + i] = s1; xp1[c * sT + i] = p1; xs2[c * sT + i] = s2; xp2[c * sT + i] = p2;
            La[c * sT + i] = 0.0;  // :1212-1215
            Ln[c * sT + i] = 0.0;  // the tail entries of both stay 0
        }
    }
    __syncwarp();

    PassIn p;
    p.sT = sT; p.T = T; p.terminated = 1;
    p.sCk = (size_t)8 * w.n_win;
    p.ck = w.ck + cb0 * p.sCk;

    for (int it = 0; it < a.n_iter; it++) {
        for (int siso = 0; siso < 2; siso++) {
            p.xs = siso == 0 ? xs1 : xs2;
            p.xp = siso == 0 ? xp1 : xp2;
            p.La = La;
            DecodeEmit emit{a, Ln, siso == 0 ? a.pi_inv : a.pi, sT, cb0, T, K, it, siso, it == a.n_iter - 1};
            siso_pass(sm, p, lane, emit);
            __syncwarp();
            double *t = La; La = Ln; Ln = t;
        }
    }
}

struct SisoEmit {
    const Ref64SisoArgs &a;
    int cb0;
    __device__ __forceinline__ int target(int) const { return 0; }
    __device__ __forceinline__ void operator()(int c, int i, int, double L, double, double) const
    {
        if (cb0 + c < a.n_cb) a.LLR[(size_t)(cb0 + c) * a.T + i] = L;
    }
};

__global__ void __launch_bounds__(32) ref64_siso_kernel(Ref64SisoArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int lane = threadIdx.x;
    const int cb0 = blockIdx.x * 4;
    if (cb0 >= a.n_cb) return;
    const int T = a.T;
    const size_t sT = T;
    const Ref64Workspace &w = a.ws;
    build_lut(sm.lut, lane);
    // de-interleave recs (xs,xp pairs) into the planar workspace
    for (int c = 0; c < 4; c++) {
        const int cb = min(cb0 + c, a.n_cb - 1);
        for (int i = lane; i < T; i += 32) {
            w.xs1[(cb0 + c) * sT + i] = a.recs[(size_t)cb * 2 * T + 2 * i];
            w.xp1[(cb0 + c) * sT + i] = a.recs[(size_t)cb * 2 * T + 2 * i + 1];
            w.La[(cb0 + c) * sT + i] = a.La[(size_t)cb * T + i];
        }
    }
    __syncwarp();
    PassIn p;
    p.xs = w.xs1 + cb0 * sT; p.xp = w.xp1 + cb0 * sT; p.La = w.La + cb0 * sT;
    p.sT = sT; p.T = T; p.terminated = a.terminated;
    p.sCk = (size_t)8 * w.n_win;
    p.ck = w.ck + cb0 * p.sCk;
    SisoEmit emit{a, cb0};
    siso_pass(sm, p, lane, emit);
}

}  // namespace

cudaError_t launch_ref64_decode(const Ref64Args &a, cudaStream_t st, int *n_launches)
{
    ref64_decode_kernel<<<(a.n_cb + 3) / 4, 32, sizeof(Smem), st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ref64_siso(const Ref64SisoArgs &a, cudaStream_t st, int *n_launches)
{
    ref64_siso_kernel<<<(a.n_cb + 3) / 4, 32, sizeof(Smem), st>>>(a);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

}  // namespace tdb200
