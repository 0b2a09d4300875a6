// tdb200_api.cu -- the C ABI declared in include/tdb200.h: plan construction (QPP tables,
// sub-block geometry), device workspace, batch chunking, host<->device staging on streams.
// No CPU fallback lives here: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "tdb200_internal.h"
#include "tdb200_plan_table.h"
#include "tdb200_plan_table_lm.h"

namespace tdb200 {

// ---------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

static int fail(int status, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

#define TDB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(e_ == cudaErrorMemoryAllocation ? TDB200_ERR_ALLOC : TDB200_ERR_CUDA,   \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

__global__ void fill_i32_kernel(int32_t *p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Every entry point works on the decoder's device and leaves the caller's current device as it found it (a thread that
// holds handles for several GPUs, or shares the process with torch, must not have its device switched under it).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define TDB_DEVICE(dev)          \
    DeviceGuard device_guard_(dev); \
    TDB_CUDA(device_guard_.err)

// ---------------------------------------------------------------------------- constants
// RSC with feedback 13_8 = 1011 and feed-forward 15_8 = 1101 (ITTC/log_map.h:34-36); state =
// 4*s0 + 2*s1 + s2 with s0 the newest register, as bin2int does (log_map.cpp:213-231).
static void rsc_step(int state, int d, int *next, int *parity)
{
    const int s0 = (state >> 2) & 1, s1 = (state >> 1) & 1, s2 = state & 1;
    const int a = (d + s1 + s2) & 1;
    *parity = (a + s0 + s2) & 1;
    *next = (a << 2) | (s0 << 1) | s1;
}

const Trellis &host_trellis()
{
    static Trellis t = [] {
        Trellis r;
        for (int s = 0; s < kStates; s++) {
            int n, p;
            rsc_step(s, 0, &n, &p);
            r.ns0[s] = n; r.par0[s] = p; r.ls0[n] = s;
            rsc_step(s, 1, &n, &p);
            r.ns1[s] = n; r.ls1[n] = s;
        }
        return r;
    }();
    return t;
}

static bool trellis_literals_ok()
{
    const Trellis &t = host_trellis();
    for (int s = 0; s < kStates; s++) {
        if (tb(kNs0, s) != t.ns0[s] || tb(kNs1, s) != t.ns1[s] || tb(kLs0, s) != t.ls0[s] ||
            tb(kLs1, s) != t.ls1[s] || (int)((kPar0 >> s) & 1) != t.par0[s])
            return false;
        int n, p;
        rsc_step(s, 1, &n, &p);
        if (p != 1 - t.par0[s]) return false;  // o1 = -o0
    }
    return true;
}

// TS 36.212 Table 5.1.3-3 (K, f1, f2).  Not present in the reference, which hard-codes
// 6144 -> (263,480) (ITTC/main.cpp:30,36-37) and mentions 2688 -> (127,504) (:17-19).
static const short kLte[188][3] = {
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

}  // namespace tdb200

using namespace tdb200;

// ---------------------------------------------------------------------------- handle
struct tdb200_decoder {
    tdb200_config cfg{};
    int T = 0, NL = 0;
    int sm_count = 0;
    int *d_pi = nullptr, *d_pi_inv = nullptr;
    std::vector<int> h_pi;
    Ref64Workspace ws64{};
    void *ws64_block = nullptr;
    FastGeom geom{};
    uint16_t *d_tab2 = nullptr;
    // staging for TDB200_MEM_HOST callers: a ring of kSlots chunks, so that the H2D copy of chunk
    // i+1, the kernel of chunk i and the D2H copy of chunk i-1 overlap (three internal streams)
    static constexpr int kSlots = 3;
    struct Slot {
        void *d_in = nullptr;
        size_t d_in_bytes = 0;
        void *d_dem = nullptr;  // demapped channel values of this chunk (tdb200_decode_symbols_batch)
        size_t d_dem_bytes = 0;
        uint8_t *d_bits = nullptr;
        int32_t *d_bits_iters = nullptr;
        int32_t *d_iters_used = nullptr;
        void *d_llr1 = nullptr, *d_llr2 = nullptr, *d_ext2 = nullptr;
        cudaEvent_t in_ready = nullptr, k_done = nullptr, out_done = nullptr;
    } slot[kSlots];
    int slot_cap = 0;  // codeblocks one slot is sized for
    cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_start = nullptr;
    void *siso_in = nullptr, *siso_out = nullptr;  // staging of tdb200_siso_batch (host callers)
    size_t siso_in_bytes = 0;
    void *dem = nullptr;  // demapped / de-rate-matched channel values, device callers of tdb200_decode_{symbols,rm}_batch
    size_t dem_bytes = 0;
    int filler = 0;  // F: filler bits at the head of every code block of this handle (tdb200_set_filler_bits)
    struct RmTable {
        int rv = 0, ncb = 0, filler = 0, nnn = 0;
        int *d_perm = nullptr, *d_inv = nullptr;
    };
    std::vector<RmTable> rm_tables;  // one per (rv, N_cb) used so far
    uint32_t *d_crc_tab = nullptr, *d_crc_shift = nullptr;  // CRC stopping rule (early_term 2 / 3)
    uint32_t crc_poly = 0;
    int launches_last = 0;
};

// Ensure a device staging buffer of at least `bytes`.
template <typename P>
static int ensure(P *&p, size_t &have, size_t bytes)
{
    if (have >= bytes) return TDB200_OK;
    cudaFree(p);
    p = nullptr; have = 0;
    TDB_CUDA(cudaMalloc(&p, bytes));
    have = bytes;
    return TDB200_OK;
}
template <typename P>
static int ensure_once(P *&p, size_t bytes)
{
    if (p) return TDB200_OK;
    TDB_CUDA(cudaMalloc(&p, bytes));
    return TDB200_OK;
}

static size_t llr_elem_size(int t) { return t == TDB200_LLR_F64 ? 8 : (t == TDB200_LLR_F32 ? 4 : (t == TDB200_LLR_F16 ? 2 : 1)); }

extern "C" {

const char *tdb200_last_error(void) { return g_err; }

const char *tdb200_status_string(int s)
{
    switch (s) {
        case TDB200_OK: return "ok";
        case TDB200_ERR_INVALID_ARG: return "invalid argument";
        case TDB200_ERR_UNSUPPORTED: return "unsupported";
        case TDB200_ERR_NO_DEVICE: return "no CUDA device";
        case TDB200_ERR_CUDA: return "CUDA error";
        case TDB200_ERR_ALLOC: return "out of memory";
        default: return "unknown status";
    }
}

int tdb200_lte_qpp_params(int K, int *f1, int *f2)
{
    for (auto &r : kLte)
        if (r[0] == K) {
            if (f1) *f1 = r[1];
            if (f2) *f2 = r[2];
            return TDB200_OK;
        }
    return fail(TDB200_ERR_INVALID_ARG, "K=%d is not an LTE turbo block size", K);
}

int tdb200_default_config(tdb200_config *cfg, int K)
{
    if (!cfg) return fail(TDB200_ERR_INVALID_ARG, "cfg is NULL");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->K = K;
    cfg->n_iter = 8;
    cfg->algo = TDB200_ALGO_MAXLOG_S16;
    return TDB200_OK;
}

void tdb200_destroy(tdb200_decoder *d)
{
    if (!d) return;
    DeviceGuard device_guard_(d->cfg.device);
    cudaFree(d->d_pi); cudaFree(d->d_pi_inv); cudaFree(d->ws64_block); cudaFree(d->d_tab2);
    cudaFree(d->siso_in); cudaFree(d->siso_out); cudaFree(d->dem);
    for (auto &t : d->rm_tables) { cudaFree(t.d_perm); cudaFree(t.d_inv); }
    cudaFree(d->d_crc_tab); cudaFree(d->d_crc_shift);
    for (auto &sl : d->slot) {
        cudaFree(sl.d_in); cudaFree(sl.d_dem); cudaFree(sl.d_bits); cudaFree(sl.d_bits_iters); cudaFree(sl.d_iters_used);
        cudaFree(sl.d_llr1); cudaFree(sl.d_llr2); cudaFree(sl.d_ext2);
        if (sl.in_ready) cudaEventDestroy(sl.in_ready);
        if (sl.k_done) cudaEventDestroy(sl.k_done);
        if (sl.out_done) cudaEventDestroy(sl.out_done);
    }
    if (d->ev_start) cudaEventDestroy(d->ev_start);
    if (d->s_h2d) cudaStreamDestroy(d->s_h2d);
    if (d->s_k) cudaStreamDestroy(d->s_k);
    if (d->s_d2h) cudaStreamDestroy(d->s_d2h);
    delete d;
}

static int create_impl(const tdb200_config *cfg, tdb200_decoder *d)
{
    if (!trellis_literals_ok()) return fail(TDB200_ERR_UNSUPPORTED, "internal: trellis literals disagree with the (13,15) generator");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(TDB200_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    }
    d->cfg = *cfg;
    tdb200_config &c = d->cfg;
    if (c.device < 0 || c.device >= ndev) return fail(TDB200_ERR_INVALID_ARG, "device %d out of range (have %d)", c.device, ndev);
    if (c.K < 8 || c.K > 8192) return fail(TDB200_ERR_INVALID_ARG, "K=%d out of range [8,8192]", c.K);
    if (c.f1 == 0 && c.f2 == 0) {
        int s = tdb200_lte_qpp_params(c.K, &c.f1, &c.f2);
        if (s != TDB200_OK) return s;
    }
    if (c.n_iter < 1 || c.n_iter > 64) return fail(TDB200_ERR_INVALID_ARG, "n_iter=%d out of range [1,64]", c.n_iter);
    d->T = c.K + kTail;
    d->NL = 3 * c.K + 4 * kTail;

    // QPP permutation pi(i) = (f1*i + f2*i^2) mod K (gen_qpp_index, log_map.cpp:616-624), built
    // with the second-difference recurrence; must be a bijection.
    const int K = c.K;
    d->h_pi.resize(K);
    std::vector<int> inv(K, -1);
    {
        long long p = 0, g = ((long long)c.f1 + c.f2) % K;  // pi(i+1) - pi(i) at i = 0
        const long long g2 = (2LL * c.f2) % K;
        for (int i = 0; i < K; i++) {
            d->h_pi[i] = (int)p;
            if (inv[p] != -1) return fail(TDB200_ERR_INVALID_ARG, "(f1=%d,f2=%d) is not a permutation for K=%d", c.f1, c.f2, K);
            inv[p] = i;
            p = (p + g) % K;
            g = (g + g2) % K;
        }
    }
    TDB_DEVICE(c.device);
    cudaDeviceProp prop;
    TDB_CUDA(cudaGetDeviceProperties(&prop, c.device));
    d->sm_count = prop.multiProcessorCount;
    TDB_CUDA(cudaMalloc(&d->d_pi, sizeof(int) * K));
    TDB_CUDA(cudaMalloc(&d->d_pi_inv, sizeof(int) * K));
    TDB_CUDA(cudaMemcpy(d->d_pi, d->h_pi.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
    TDB_CUDA(cudaMemcpy(d->d_pi_inv, inv.data(), sizeof(int) * K, cudaMemcpyHostToDevice));

    if (c.algo == TDB200_ALGO_LOGMAP_F64) {
        if (c.early_term) return fail(TDB200_ERR_UNSUPPORTED, "early termination is not part of the reference-order fp64 mode");
        if (c.max_batch <= 0) c.max_batch = 4096;
        const size_t nb = (size_t)((c.max_batch + 3) / 4) * 4;  // warps own groups of four
        const size_t T = d->T;
        const size_t n_win = (T + kRef64Window - 1) / kRef64Window;
        const size_t per_cb = 6 * T + 8 * n_win;
        TDB_CUDA(cudaMalloc(&d->ws64_block, sizeof(double) * per_cb * nb));
        double *p = static_cast<double *>(d->ws64_block);
        Ref64Workspace &w = d->ws64;
        w.xs1 = p; p += T * nb;
        w.xp1 = p; p += T * nb;
        w.xs2 = p; p += T * nb;
        w.xp2 = p; p += T * nb;
        w.La = p; p += T * nb;
        w.Le = p; p += T * nb;
        w.ck = p; p += 8 * n_win * nb;
        w.n_win = (int)n_win;
        w.max_batch = c.max_batch;
    } else if (c.algo == TDB200_ALGO_MAXLOG_S16 || c.algo == TDB200_ALGO_LOGMAP_F32 || c.algo == TDB200_ALGO_MAXLOG_F32 ||
               c.algo == TDB200_ALGO_LINLOGMAP_F32 || c.algo == TDB200_ALGO_LOGMAP_S16) {
        const bool lm16 = (c.algo == TDB200_ALGO_LOGMAP_S16);
        const bool s16 = (c.algo == TDB200_ALGO_MAXLOG_S16) || lm16;
        if (c.max_batch <= 0) c.max_batch = 16384;
        if (c.early_term < 0 || c.early_term > 3) return fail(TDB200_ERR_INVALID_ARG, "early_term=%d (0..3)", c.early_term);
        if (c.early_term >= 2 && !s16) return fail(TDB200_ERR_UNSUPPORTED, "the CRC stopping rule exists in the packed 16-bit decoders only (TDB200_ALGO_MAXLOG_S16, TDB200_ALGO_LOGMAP_S16)");
        if (c.early_term >= 2 && K <= 24) return fail(TDB200_ERR_INVALID_ARG, "K=%d leaves no room for a 24-bit CRC", K);
        if (c.frac_bits == 0) c.frac_bits = lm16 ? 4 : 3;
        if (lm16 && c.frac_bits < 3) return fail(TDB200_ERR_INVALID_ARG, "frac_bits=%d: TDB200_ALGO_LOGMAP_S16 needs 3 or 4 (the correction is 5 << (frac_bits - 3) at most)", c.frac_bits);
        if (c.frac_bits < 1 || c.frac_bits > 4) return fail(TDB200_ERR_INVALID_ARG, "frac_bits=%d out of range [1,4]", c.frac_bits);
        if (c.ext_scale_q2 == 0) c.ext_scale_q2 = (c.algo == TDB200_ALGO_LOGMAP_F32 || c.algo == TDB200_ALGO_LINLOGMAP_F32 || lm16) ? 4 : 3;
        if (c.ext_scale_q2 != 3 && c.ext_scale_q2 != 4) return fail(TDB200_ERR_INVALID_ARG, "ext_scale_q2=%d (3 or 4)", c.ext_scale_q2);
        if (c.et_threshold == 0) c.et_threshold = 1 << (c.frac_bits + 3);
        if (c.et_threshold < 1 || c.et_threshold > 4096 || (c.et_threshold & (c.et_threshold - 1)))
            return fail(TDB200_ERR_INVALID_ARG, "et_threshold=%d must be a power of two in [1,4096]", c.et_threshold);
        if (c.ext_clip == 0) c.ext_clip = lm16 ? 511 : (1 << (c.frac_bits + 6)) - 1;  // |Le| < 64.0 (32.0 at 4 fractional bits): keeps every metric sum inside int16
        if (c.ext_clip < 63 || c.ext_clip > 2047 || ((c.ext_clip + 1) & 3))
            return fail(TDB200_ERR_INVALID_ARG, "ext_clip=%d: need 63 <= ext_clip <= 2047 and ext_clip+1 a multiple of 4", c.ext_clip);
        // ---- sub-block geometry: K = P * L, L = 8 * NW, P <= 256 threads
        FastGeom &g = d->geom;
        int L = c.sub_block;
        if (s16 && L == 0 && c.warmup == 0) {
            // measured on a B200 for every LTE block size (tools/tune_subblock.py): which admissible L is fastest
            // depends on how the sub-block count fills warps and how many CTAs fit an SM, not on L alone -- and it is
            // not the same L for the Log-MAP kernels (245 registers, one more guard window): they have their own table
            const unsigned char *tuned = lm16 ? kTunedL8Lm : kTunedL8;
            for (int i = 0; i < 188; i++)
                if (kLte[i][0] == K && tuned[i]) { L = 8 * tuned[i]; break; }
            if (L && (K % L || K / L > 256)) L = 0;
        }
        if (L == 0 && K <= 56) L = K;  // shortest blocks: one thread walks the whole trellis (shared memory admits >= 256 threads per SM only for L <= 56)
        if (L == 0) {
            static const int pref[] = {48, 40, 56, 32, 64, 24, 72, 80, 96, 16, 128, 8};
            for (int cand : pref)
                if (K % cand == 0 && K / cand <= 256) { L = cand; break; }
            if (L == 0)
                for (int cand = 8; cand <= K; cand += 8)
                    if (K % cand == 0 && K / cand <= 256) { L = cand; break; }
        }
        if (L < 8 || L % 8 || K % L || K / L > 256)
            return fail(TDB200_ERR_INVALID_ARG, "sub_block=%d must be a multiple of 8 dividing K=%d with K/sub_block <= 256", L, K);
        int G = c.warmup;
        if (G < 0 || G % 8) return fail(TDB200_ERR_INVALID_ARG, "warmup=%d must be a non-negative multiple of 8", G);
        if (c.warmup == 0 && c.sub_block == 0) G = lm16 ? 24 : 16;  // auto plan: guard of 16 (BER vs (L,G)); Log-MAP: 24, from where on the block-error rate no longer moves
        // A guard of two sub-blocks exists for 8-step sub-blocks only (the block sizes K = 8 x prime, where nothing but 8
        // and K divides K): with guard 8 those plans lose 0.06-0.09 dB against the unsegmented recursion
        // (profiles/r02_plan_ber_parity_*.json); the kernels with run-time geometry walk the guard across two neighbours.
        if (c.warmup == 0 && c.sub_block == 0 && L == 8 && K / L > 2) G = 16;
        if (G > L && !(L == 8 && G == 16 && K / L > 2)) G = L;
        g.K = K; g.L = L; g.P = K / L; g.NW = L / 8; g.G = (g.P == 1) ? 0 : G;
        g.PP = g.P | 1;  // odd row pitch: de-multiplex stores spread over the banks, walks stay conflict-free
        g.threads = ((g.P + 31) / 32) * 32;
        g.n_ckpt = std::max(g.NW - 2, 0);
        g.NP = 1; g.pair_bytes = 0;
        if (s16) {
            g.pair_bytes = fast_s16_pair_bytes(g);
            // Pairs per CTA.  Fewer than 32 sub-blocks: fill two warps (measured: larger CTAs only add barrier
            // coupling -- registers cap an SM at eight or nine warps either way).  33..42 sub-blocks leave most of
            // the CTA's second warp idle: three pairs share a full 128-thread CTA instead (measured on B200,
            // tools/tune_pairs.py, profiles/r01_pairs_tuning.json: +40 % at L <= 32, +4 % at L = 64; two pairs in a
            // 96-thread CTA gain nothing -- three-warp CTAs load the four sub-partitions unevenly).  49..64
            // sub-blocks: two pairs per CTA only where shared memory admits three single-pair CTAs per SM but two
            // doubles (K = 3136: +13 %).
            int np = 1;
            if (c.early_term < 2) {  // the CRC fold is per CTA
                const bool packable = lm16 ? !fast_s16_specialised(g, true) : (!fast_s16_specialised(g) || fast_spec_rt(g));
                auto ctas_by_smem = [&](int n) {
                    g.NP = n; g.threads = ((n * g.P + 31) / 32) * 32;
                    return (int)(prop.sharedMemPerMultiprocessor / ((size_t)fast_s16_smem_bytes(g) + 1024));
                };
                if (g.P < 32) np = std::max(1, 64 / g.P);
                else if (packable && g.P > 32 && 3 * g.P <= 128) np = 3;
                else if (packable && 2 * g.P <= 128 && 2 * g.P > 96 && ctas_by_smem(1) == 3 && ctas_by_smem(2) >= 2) np = 2;
                if (const char *e = getenv("TDB200_PAIRS_PER_CTA")) {  // tuning knob (tools/tune_pairs.py)
                    const int want = atoi(e);
                    if (want >= 1 && (want == 1 || (packable && want * g.P <= 128))) np = want;
                }
            }
            while (np > 1) {
                g.NP = np; g.threads = ((np * g.P + 31) / 32) * 32;
                if ((size_t)fast_s16_smem_bytes(g) <= prop.sharedMemPerBlockOptin) break;
                np--;
            }
            g.NP = np; g.threads = ((np * g.P + 31) / 32) * 32;
        }
        g.smem_bytes = s16 ? fast_s16_smem_bytes(g) : f32_smem_bytes(g);
        if ((size_t)g.smem_bytes > prop.sharedMemPerBlockOptin)
            return fail(TDB200_ERR_UNSUPPORTED, "plan needs %d B of shared memory per CTA, device allows %zu", g.smem_bytes, (size_t)prop.sharedMemPerBlockOptin);
        c.sub_block = L; c.warmup = g.G;
        if (s16) TDB_CUDA(fast_s16_configure(g, d->sm_count, lm16));
        else TDB_CUDA(f32_configure(g));
        // word address of element pi(tL+j), stored at j*PP+t
        std::vector<uint16_t> tab((size_t)L * g.PP, 0);
        for (int i = 0; i < K; i++) {
            const int t = i / L, j = i % L, n = d->h_pi[i];
            tab[j * g.PP + t] = (uint16_t)((n % L) * g.PP + n / L);
        }
        if (c.early_term >= 2) {
            // byte-wise table of the generator and the weight x^((P-1-t)L) mod g of sub-block t's remainder
            const uint32_t poly = d->crc_poly = (c.early_term == 2) ? 0x800063u : 0x864CFBu;
            std::vector<uint32_t> tb(256), sh(g.P);
            for (uint32_t v = 0; v < 256; v++) {
                uint32_t r = v << 16;
                for (int b = 0; b < 8; b++) r = ((r << 1) & 0xffffffu) ^ ((r & 0x800000u) ? poly : 0u);
                tb[v] = r;
            }
            uint32_t a = 1;
            for (int t = g.P - 1; t >= 0; t--) {
                sh[t] = a;
                for (int b = 0; b < L; b++) a = ((a << 1) & 0xffffffu) ^ ((a & 0x800000u) ? poly : 0u);
            }
            TDB_CUDA(cudaMalloc(&d->d_crc_tab, sizeof(uint32_t) * 256));
            TDB_CUDA(cudaMalloc(&d->d_crc_shift, sizeof(uint32_t) * g.P));
            TDB_CUDA(cudaMemcpy(d->d_crc_tab, tb.data(), sizeof(uint32_t) * 256, cudaMemcpyHostToDevice));
            TDB_CUDA(cudaMemcpy(d->d_crc_shift, sh.data(), sizeof(uint32_t) * g.P, cudaMemcpyHostToDevice));
        }
        TDB_CUDA(cudaMalloc(&d->d_tab2, sizeof(uint16_t) * tab.size()));
        TDB_CUDA(cudaMemcpy(d->d_tab2, tab.data(), sizeof(uint16_t) * tab.size(), cudaMemcpyHostToDevice));
    } else {
        return fail(TDB200_ERR_INVALID_ARG, "algo=%d", c.algo);
    }
    return TDB200_OK;
}

int tdb200_create(const tdb200_config *cfg, tdb200_decoder **out)
{
    if (!cfg || !out) return fail(TDB200_ERR_INVALID_ARG, "cfg/out is NULL");
    *out = nullptr;
    tdb200_decoder *d = new (std::nothrow) tdb200_decoder;
    if (!d) return fail(TDB200_ERR_ALLOC, "host allocation failed");
    int s = create_impl(cfg, d);
    if (s != TDB200_OK) {
        tdb200_destroy(d);
        return s;
    }
    *out = d;
    return TDB200_OK;
}

int tdb200_get_plan(const tdb200_decoder *d, tdb200_plan_info *info)
{
    if (!d || !info) return fail(TDB200_ERR_INVALID_ARG, "dec/info is NULL");
    std::memset(info, 0, sizeof(*info));
    info->K = d->cfg.K; info->f1 = d->cfg.f1; info->f2 = d->cfg.f2;
    info->n_iter = d->cfg.n_iter; info->algo = d->cfg.algo;
    info->max_batch = d->cfg.max_batch; info->sm_count = d->sm_count;
    info->kernel_launches_last_call = d->launches_last;
    if (d->cfg.algo == TDB200_ALGO_LOGMAP_F64) {
        info->sub_block = d->cfg.K; info->n_sub_blocks = 1; info->cb_per_cta = 8; info->threads_per_cta = 64;
    } else {
        info->sub_block = d->geom.L; info->n_sub_blocks = d->geom.P; info->warmup = d->geom.G;
        info->cb_per_cta = (d->cfg.algo == TDB200_ALGO_MAXLOG_S16 || d->cfg.algo == TDB200_ALGO_LOGMAP_S16) ? 2 * d->geom.NP : 1;
        info->threads_per_cta = d->geom.threads; info->smem_bytes = d->geom.smem_bytes;
    }
    return TDB200_OK;
}

// Lazily create the pipeline objects of a handle (first host-memory call).
static int ensure_pipeline(tdb200_decoder *d)
{
    if (d->s_h2d) return TDB200_OK;
    TDB_CUDA(cudaStreamCreateWithFlags(&d->s_h2d, cudaStreamNonBlocking));
    TDB_CUDA(cudaStreamCreateWithFlags(&d->s_k, cudaStreamNonBlocking));
    TDB_CUDA(cudaStreamCreateWithFlags(&d->s_d2h, cudaStreamNonBlocking));
    TDB_CUDA(cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming));
    for (auto &sl : d->slot) {
        TDB_CUDA(cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming));
        TDB_CUDA(cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming));
        TDB_CUDA(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
    }
    return TDB200_OK;
}

// One chunk of n codeblocks, all pointers device pointers, enqueued on st.
static int launch_chunk(tdb200_decoder *d, const void *v_llr, int llr_type, int n, uint8_t *v_bits, int32_t *v_bits_iters,
                        int32_t *v_iters, void *v_llr1, void *v_llr2, void *v_ext2, cudaStream_t st, const void *v_sym_q = nullptr, double kf = 0.0)
{
    const tdb200_config &c = d->cfg;
    if (c.algo == TDB200_ALGO_LOGMAP_F64) {
        Ref64Args a{};
        a.llr = v_llr; a.llr_type = llr_type; a.n_cb = n; a.K = c.K; a.n_iter = c.n_iter;
        a.pi = d->d_pi; a.pi_inv = d->d_pi_inv; a.ws = d->ws64;
        a.bits = v_bits; a.bits_iters = v_bits_iters;
        a.llr1 = static_cast<double *>(v_llr1); a.llr2 = static_cast<double *>(v_llr2); a.ext2 = static_cast<double *>(v_ext2);
        TDB_CUDA(launch_ref64_decode(a, st, &d->launches_last));
    } else if (c.algo != TDB200_ALGO_MAXLOG_S16 && c.algo != TDB200_ALGO_LOGMAP_S16) {
        F32Args a{};
        a.llr = v_llr; a.llr_type = llr_type; a.n_cb = n; a.g = d->geom; a.n_iter = c.n_iter;
        a.logmap = (c.algo == TDB200_ALGO_LOGMAP_F32) ? 1 : (c.algo == TDB200_ALGO_LINLOGMAP_F32 ? 2 : 0);
        a.ext_scale = 0.25f * (float)c.ext_scale_q2;
        a.ext_clamp = 1.0e6f;  // the fp32 modes do not clamp the extrinsic (the reference does not either)
        a.early_term = c.early_term;
        a.et_threshold = (float)c.et_threshold / (float)(1 << c.frac_bits);
        a.tab2 = d->d_tab2;
        a.bits = v_bits; a.iters_used = v_iters;
        a.llr2 = static_cast<float *>(v_llr2); a.ext2 = static_cast<float *>(v_ext2);
        TDB_CUDA(launch_f32(a, st, &d->launches_last));
    } else {
        FastArgs a{};
        a.llr = v_llr; a.llr_type = llr_type; a.n_cb = n; a.g = d->geom; a.n_iter = c.n_iter;
        a.sym_q = v_sym_q; a.kf = (float)kf;
        a.frac_bits = c.frac_bits;
        a.llr_clip = std::min((1 << (c.frac_bits + 4)) - 1, 127);  // systematic values are kept as bytes in shared memory
        a.ext_lim = c.ext_clip + 1;
        a.q2 = c.ext_scale_q2; a.early_term = std::min(c.early_term, 2); a.et_threshold = c.et_threshold;
        a.crc_poly = d->crc_poly; a.crc_tab = d->d_crc_tab; a.crc_shift = d->d_crc_shift;
        a.logmap = (c.algo == TDB200_ALGO_LOGMAP_S16) ? 1 : 0;
        a.lm_t4 = 5 << (c.frac_bits - 3 > 0 ? c.frac_bits - 3 : 0);
        a.tab2 = d->d_tab2;
        a.opaque[0] = 0xffffffffu; a.opaque[1] = 4u; a.opaque[2] = 65536u; a.opaque[3] = 0xC0000000u;
        a.prefetch_stride = d->geom.resident_ctas;
        a.sm_count = d->sm_count;
        a.bits = v_bits; a.bits_iters = v_bits_iters; a.iters_used = v_iters;
        a.llr2 = static_cast<float *>(v_llr2); a.ext2 = static_cast<float *>(v_ext2);
        TDB_CUDA(launch_fast_s16(a, st, &d->launches_last));
    }
    return TDB200_OK;
}

// Where the channel values of a decode call come from: LLRs as they are, or received symbols that
// are demapped on the device into the decoder's own input format first.
struct Source {
    const void *llr = nullptr;
    int llr_type = 0;
    const void *sym_i = nullptr, *sym_q = nullptr;
    int sym_type = 0, modulation = 0;
    double kf = 0.0;
    bool symbols() const { return sym_i != nullptr; }
    // rate-matched LLRs: `llr` holds [n_cb][rm_E] values that are de-rate-matched on the device first
    const tdb200_decoder::RmTable *rm = nullptr;
    int rm_E = 0;
};

// the channel-value format the decoder consumes without conversion loss
static int native_llr_type(const tdb200_config &c)
{
    return c.algo == TDB200_ALGO_LOGMAP_F64 ? TDB200_LLR_F64 : ((c.algo == TDB200_ALGO_MAXLOG_S16 || c.algo == TDB200_ALGO_LOGMAP_S16) ? TDB200_LLR_S8 : TDB200_LLR_F32);
}

static int launch_demap_chunk(tdb200_decoder *d, const Source &src, const void *si, const void *sq, void *llr, int n, cudaStream_t st)
{
    DemapArgs a{};
    a.sym_i = si; a.sym_q = sq; a.sym_type = src.sym_type;
    a.llr = llr; a.llr_type = native_llr_type(d->cfg);
    a.n_llr = (size_t)n * d->NL; a.modulation = src.modulation; a.kf = src.kf;
    a.frac_bits = d->cfg.frac_bits; a.clip = std::min((1 << (d->cfg.frac_bits + 4)) - 1, 127);
    TDB_CUDA(launch_demap(a, st));
    d->launches_last += 1;
    return TDB200_OK;
}

static int launch_dematch_chunk(tdb200_decoder *d, const Source &src, const void *e_llr, void *llr, int out_type, int n, cudaStream_t st)
{
    RmArgs a{};
    a.e_llr = e_llr; a.llr = llr; a.in_type = src.llr_type; a.out_type = out_type;
    a.inv = src.rm->d_inv; a.nnn = src.rm->nnn; a.NL = d->NL; a.E = src.rm_E; a.n_cb = n; a.accumulate = 0;
    a.frac_bits = d->cfg.frac_bits ? d->cfg.frac_bits : 3;
    a.clip = std::min((1 << (a.frac_bits + 4)) - 1, 127);
    TDB_CUDA(launch_rate_dematch(a, st));
    d->launches_last += 1;
    return TDB200_OK;
}

int tdb200_set_filler_bits(tdb200_decoder *d, int F)
{
    if (!d) return fail(TDB200_ERR_INVALID_ARG, "dec is NULL");
    if (F < 0 || F >= d->cfg.K) return fail(TDB200_ERR_INVALID_ARG, "F=%d filler bits (0 <= F < K=%d)", F, d->cfg.K);
    d->filler = F;
    return TDB200_OK;
}

// The (rv, N_cb, F) permutation of a handle, built on first use.
static int rm_table(tdb200_decoder *d, int rv, int ncb, const tdb200_decoder::RmTable **out)
{
    if (rv < 0 || rv > 3) return fail(TDB200_ERR_INVALID_ARG, "rv=%d (0..3)", rv);
    const int Kw = 3 * 32 * ((d->cfg.K + 4 + 31) / 32);
    if (ncb < 0) return fail(TDB200_ERR_INVALID_ARG, "ncb=%d", ncb);
    if (ncb == 0 || ncb > Kw) ncb = Kw;
    for (auto &t : d->rm_tables)
        if (t.rv == rv && t.ncb == ncb && t.filler == d->filler) { *out = &t; return TDB200_OK; }
    std::vector<int> perm, inv;
    if (!build_rm_table(d->cfg.K, rv, ncb, d->filler, perm, inv)) return fail(TDB200_ERR_INVALID_ARG, "ncb=%d leaves no transmittable bit", ncb);
    tdb200_decoder::RmTable t;
    t.rv = rv; t.ncb = ncb; t.filler = d->filler; t.nnn = (int)perm.size();
    TDB_DEVICE(d->cfg.device);
    TDB_CUDA(cudaMalloc(&t.d_perm, sizeof(int) * perm.size()));
    if (cudaMalloc(&t.d_inv, sizeof(int) * inv.size()) != cudaSuccess) { cudaFree(t.d_perm); return fail(TDB200_ERR_ALLOC, "device allocation failed"); }
    cudaError_t e = cudaMemcpy(t.d_perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_inv, inv.data(), sizeof(int) * inv.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(t.d_perm); cudaFree(t.d_inv); return fail(TDB200_ERR_CUDA, "rate-matching table upload: %s", cudaGetErrorString(e)); }
    d->rm_tables.reserve(64);  // pointers into the vector are handed out: keep them stable
    if (d->rm_tables.size() >= 64) { cudaFree(t.d_perm); cudaFree(t.d_inv); return fail(TDB200_ERR_UNSUPPORTED, "more than 64 distinct (rv, ncb, filler) triples on one handle"); }
    d->rm_tables.push_back(t);
    *out = &d->rm_tables.back();
    return TDB200_OK;
}

static int decode_core(tdb200_decoder *d, const Source &src, int mem, int n_cb, const tdb200_outputs *out, void *stream)
{
    if (n_cb < 0) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d", n_cb);
    if (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE) return fail(TDB200_ERR_INVALID_ARG, "mem=%d", mem);
    d->launches_last = 0;
    if (n_cb == 0) return TDB200_OK;
    const tdb200_config &c = d->cfg;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(c.device);
    const int K = c.K, T = d->T, NL = d->NL;
    const bool sym = src.symbols(), rm = (src.rm != nullptr);
    if (mem == TDB200_MEM_DEVICE && !sym && !rm) {
        // the load stage reads 12 values at a time with the widest loads the row pitch allows
        const uintptr_t need = src.llr_type == TDB200_LLR_S8 ? 4 : (src.llr_type == TDB200_LLR_F16 ? 8 : 16);
        if (reinterpret_cast<uintptr_t>(src.llr) & (need - 1))
            return fail(TDB200_ERR_INVALID_ARG, "device LLR buffer must be %d-byte aligned", (int)need);
    }
    // what the decode kernel reads: the caller's LLRs, or the staged output of the demapper / de-rate-matcher
    const int llr_type = sym ? native_llr_type(c) : ((rm && (c.algo == TDB200_ALGO_MAXLOG_S16 || c.algo == TDB200_ALGO_LOGMAP_S16)) ? TDB200_LLR_S8 : src.llr_type);
    const void *llr = src.llr;
    const size_t esz = llr_elem_size(llr_type);
    const size_t ssz = sym ? llr_elem_size(src.sym_type) : 0;
    const size_t NS = sym ? (size_t)NL / src.modulation : 0;  // symbols per codeblock
    const size_t rsz = rm ? llr_elem_size(src.llr_type) : 0, RE = rm ? (size_t)src.rm_E : 0;  // rate-matched row
    // BPSK / QPSK float symbols into a packed decoder: the demapper runs inside the decoder's load stage (no staging
    // buffer, no demap kernel); every other combination demaps into the decoder's native format first
    const bool fused_sym = sym && (c.algo == TDB200_ALGO_MAXLOG_S16 || c.algo == TDB200_ALGO_LOGMAP_S16) && c.early_term < 2 &&
                           src.sym_type == TDB200_LLR_F32 && (src.modulation == 1 || src.modulation == 2) && !getenv("TDB200_NO_FUSED_DEMAP") &&
                           (mem == TDB200_MEM_HOST ||  // staged planes are aligned; caller's device planes must allow the 128- / 64-bit loads
                            ((reinterpret_cast<uintptr_t>(src.sym_i) & (src.modulation == 1 ? 15 : 7)) == 0 &&
                             (src.modulation == 1 || (reinterpret_cast<uintptr_t>(src.sym_q) & 7) == 0)));
    const int fused_type = src.modulation == 1 ? kLlrSymBpskF32 : kLlrSymQpskF32;
    const bool staged = (sym && !fused_sym) || rm;

    const bool f64 = (c.algo == TDB200_ALGO_LOGMAP_F64);
    const size_t fsz = f64 ? sizeof(double) : sizeof(float);  // native float type of the LLR outputs
    const bool s16 = (c.algo == TDB200_ALGO_MAXLOG_S16 || c.algo == TDB200_ALGO_LOGMAP_S16);
    if (!f64 && out->llr_siso1)
        return fail(TDB200_ERR_UNSUPPORTED, "llr_siso1 is produced by TDB200_ALGO_LOGMAP_F64 only");
    if (!f64 && !s16 && out->bits_iters)
        return fail(TDB200_ERR_UNSUPPORTED, "bits_iters is produced by the fp64 and the packed 16-bit decoders only");
    if (!f64 && c.early_term >= 2 && out->ext_siso2)
        return fail(TDB200_ERR_UNSUPPORTED, "ext_siso2 is not available with the CRC stopping rule (a block may stop after SISO-1)");
    if (!f64 && c.early_term && out->llr_siso2)
        return fail(TDB200_ERR_UNSUPPORTED, "llr_siso2 is not available with early termination (the stopping iteration is not known in advance)");

    if (mem == TDB200_MEM_DEVICE) {
        // ---- device buffers: chunks of max_batch (the fp64 workspace is sized for that), all on `stream`
        if (staged) {
            int s = ensure(d->dem, d->dem_bytes, (size_t)std::min(c.max_batch, n_cb) * NL * esz);
            if (s) return s;
        }
        for (int c0 = 0; c0 < n_cb; c0 += c.max_batch) {
            const int n = std::min(c.max_batch, n_cb - c0);
            if (sym && !fused_sym) {
                int s = launch_demap_chunk(d, src, static_cast<const char *>(src.sym_i) + (size_t)c0 * NS * ssz,
                                           static_cast<const char *>(src.sym_q) + (size_t)c0 * NS * ssz, d->dem, n, st);
                if (s) return s;
            } else if (rm) {
                int s = launch_dematch_chunk(d, src, static_cast<const char *>(llr) + (size_t)c0 * RE * rsz, d->dem, llr_type, n, st);
                if (s) return s;
            }
            const void *in = staged ? d->dem : (fused_sym ? static_cast<const char *>(src.sym_i) + (size_t)c0 * NS * ssz
                                                          : static_cast<const char *>(llr) + (size_t)c0 * NL * esz);
            int s = launch_chunk(d, in, fused_sym ? fused_type : llr_type, n,
                                 out->bits ? out->bits + (size_t)c0 * K : nullptr,
                                 out->bits_iters ? out->bits_iters + (size_t)c0 * c.n_iter * K : nullptr,
                                 out->iters_used ? out->iters_used + c0 : nullptr,
                                 out->llr_siso1 ? static_cast<char *>(out->llr_siso1) + fsz * (size_t)c0 * T : nullptr,
                                 out->llr_siso2 ? static_cast<char *>(out->llr_siso2) + fsz * (size_t)c0 * T : nullptr,
                                 out->ext_siso2 ? static_cast<char *>(out->ext_siso2) + fsz * (size_t)c0 * T : nullptr, st,
                                 fused_sym ? static_cast<const char *>(src.sym_q) + (size_t)c0 * NS * ssz : nullptr, src.kf);
            if (s) return s;
        }
        if (f64 && out->iters_used) {  // the fp64 mode always runs every iteration: a fill kernel, so that the call stays asynchronous
            fill_i32_kernel<<<(n_cb + 255) / 256, 256, 0, st>>>(out->iters_used, n_cb, c.n_iter);
            TDB_CUDA(cudaGetLastError());
        }
        return TDB200_OK;
    }

    // ---- host buffers: a three-stage pipeline over chunks of `cap` codeblocks.  With pinned host
    //      memory the H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1
    //      run concurrently; with pageable memory the copies serialise but the result is the same.
    int s = ensure_pipeline(d);
    if (s) return s;
    if (d->slot_cap == 0) {
        // about eight chunks per max_batch, at least one full wave of CTAs for the throughput kernel
        int cap = std::max(64, c.max_batch / 8);
        if (!f64) cap = std::max(cap, 2 * std::max(d->geom.resident_ctas, 1));
        d->slot_cap = std::min(cap, c.max_batch);
    }
    const int cap = d->slot_cap;
    for (auto &sl : d->slot) {
        if ((s = ensure(sl.d_in, sl.d_in_bytes, sym ? 2 * (size_t)cap * NS * ssz : (rm ? (size_t)cap * RE * rsz : (size_t)cap * NL * esz)))) return s;
        if (staged && (s = ensure(sl.d_dem, sl.d_dem_bytes, (size_t)cap * NL * esz))) return s;
        if (out->bits && (s = ensure_once(sl.d_bits, (size_t)cap * K))) return s;
        if (out->bits_iters && (s = ensure_once(sl.d_bits_iters, sizeof(int32_t) * (size_t)cap * c.n_iter * K))) return s;
        if (out->iters_used && !f64 && (s = ensure_once(sl.d_iters_used, sizeof(int32_t) * (size_t)cap))) return s;
        if (out->llr_siso1 && (s = ensure_once(sl.d_llr1, 8 * (size_t)cap * T))) return s;
        if (out->llr_siso2 && (s = ensure_once(sl.d_llr2, 8 * (size_t)cap * T))) return s;
        if (out->ext_siso2 && (s = ensure_once(sl.d_ext2, 8 * (size_t)cap * T))) return s;
    }
    // work already queued on the caller's stream comes first
    TDB_CUDA(cudaEventRecord(d->ev_start, st));
    TDB_CUDA(cudaStreamWaitEvent(d->s_h2d, d->ev_start, 0));
    TDB_CUDA(cudaStreamWaitEvent(d->s_k, d->ev_start, 0));
    TDB_CUDA(cudaStreamWaitEvent(d->s_d2h, d->ev_start, 0));
    int i = 0;
    for (int c0 = 0; c0 < n_cb; c0 += cap, i++) {
        const int n = std::min(cap, n_cb - c0);
        tdb200_decoder::Slot &sl = d->slot[i % tdb200_decoder::kSlots];
        const bool reused = i >= tdb200_decoder::kSlots;
        if (reused) TDB_CUDA(cudaStreamWaitEvent(d->s_h2d, sl.k_done, 0));  // the kernel that read this slot's input
        char *d_q = static_cast<char *>(sl.d_in) + (size_t)cap * NS * ssz;  // the slot's Q plane
        if (sym) {
            TDB_CUDA(cudaMemcpyAsync(sl.d_in, static_cast<const char *>(src.sym_i) + (size_t)c0 * NS * ssz, (size_t)n * NS * ssz,
                                     cudaMemcpyHostToDevice, d->s_h2d));
            TDB_CUDA(cudaMemcpyAsync(d_q, static_cast<const char *>(src.sym_q) + (size_t)c0 * NS * ssz, (size_t)n * NS * ssz,
                                     cudaMemcpyHostToDevice, d->s_h2d));
        } else if (rm) {
            TDB_CUDA(cudaMemcpyAsync(sl.d_in, static_cast<const char *>(llr) + (size_t)c0 * RE * rsz, (size_t)n * RE * rsz,
                                     cudaMemcpyHostToDevice, d->s_h2d));
        } else {
            TDB_CUDA(cudaMemcpyAsync(sl.d_in, static_cast<const char *>(llr) + (size_t)c0 * NL * esz, (size_t)n * NL * esz,
                                     cudaMemcpyHostToDevice, d->s_h2d));
        }
        TDB_CUDA(cudaEventRecord(sl.in_ready, d->s_h2d));
        TDB_CUDA(cudaStreamWaitEvent(d->s_k, sl.in_ready, 0));
        if (reused) TDB_CUDA(cudaStreamWaitEvent(d->s_k, sl.out_done, 0));  // the copy-out of this slot's previous results
        if (sym && !fused_sym && (s = launch_demap_chunk(d, src, sl.d_in, d_q, sl.d_dem, n, d->s_k))) return s;
        if (rm && (s = launch_dematch_chunk(d, src, sl.d_in, sl.d_dem, llr_type, n, d->s_k))) return s;
        s = launch_chunk(d, staged ? sl.d_dem : sl.d_in, fused_sym ? fused_type : llr_type, n, out->bits ? sl.d_bits : nullptr,
                         out->bits_iters ? sl.d_bits_iters : nullptr,
                         (out->iters_used && !f64) ? sl.d_iters_used : nullptr, out->llr_siso1 ? sl.d_llr1 : nullptr,
                         out->llr_siso2 ? sl.d_llr2 : nullptr, out->ext_siso2 ? sl.d_ext2 : nullptr, d->s_k, fused_sym ? d_q : nullptr, src.kf);
        if (s) return s;
        TDB_CUDA(cudaEventRecord(sl.k_done, d->s_k));
        TDB_CUDA(cudaStreamWaitEvent(d->s_d2h, sl.k_done, 0));
        cudaStream_t so = d->s_d2h;
        if (out->bits) TDB_CUDA(cudaMemcpyAsync(out->bits + (size_t)c0 * K, sl.d_bits, (size_t)n * K, cudaMemcpyDeviceToHost, so));
        if (out->bits_iters) TDB_CUDA(cudaMemcpyAsync(out->bits_iters + (size_t)c0 * c.n_iter * K, sl.d_bits_iters, sizeof(int32_t) * (size_t)n * c.n_iter * K, cudaMemcpyDeviceToHost, so));
        if (out->iters_used && !f64) TDB_CUDA(cudaMemcpyAsync(out->iters_used + c0, sl.d_iters_used, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, so));
        if (out->llr_siso1) TDB_CUDA(cudaMemcpyAsync(static_cast<char *>(out->llr_siso1) + fsz * (size_t)c0 * T, sl.d_llr1, fsz * (size_t)n * T, cudaMemcpyDeviceToHost, so));
        if (out->llr_siso2) TDB_CUDA(cudaMemcpyAsync(static_cast<char *>(out->llr_siso2) + fsz * (size_t)c0 * T, sl.d_llr2, fsz * (size_t)n * T, cudaMemcpyDeviceToHost, so));
        if (out->ext_siso2) TDB_CUDA(cudaMemcpyAsync(static_cast<char *>(out->ext_siso2) + fsz * (size_t)c0 * T, sl.d_ext2, fsz * (size_t)n * T, cudaMemcpyDeviceToHost, so));
        TDB_CUDA(cudaEventRecord(sl.out_done, so));
    }
    // results are in the caller's memory when this returns; the caller's stream is ordered after them
    for (int k = 0; k < std::min(i, (int)tdb200_decoder::kSlots); k++) TDB_CUDA(cudaStreamWaitEvent(st, d->slot[k].out_done, 0));
    TDB_CUDA(cudaStreamSynchronize(d->s_d2h));
    if (f64 && out->iters_used)
        for (int k = 0; k < n_cb; k++) out->iters_used[k] = c.n_iter;  // the fp64 mode always runs every iteration
    return TDB200_OK;
}

int tdb200_decode_batch(tdb200_decoder *d, const void *llr, int llr_type, int mem, int n_cb,
                        const tdb200_outputs *out, void *stream)
{
    if (!d || !llr || !out) return fail(TDB200_ERR_INVALID_ARG, "dec/llr/out is NULL");
    if (llr_type < TDB200_LLR_F64 || llr_type > TDB200_LLR_F16) return fail(TDB200_ERR_INVALID_ARG, "llr_type=%d", llr_type);
    Source src;
    src.llr = llr; src.llr_type = llr_type;
    return decode_core(d, src, mem, n_cb, out, stream);
}

static int check_sym(const char *fn, const tdb200_decoder *d, int sym_type, int modulation)
{
    if (sym_type != TDB200_LLR_F32 && sym_type != TDB200_LLR_F64 && sym_type != TDB200_LLR_F16)
        return fail(TDB200_ERR_INVALID_ARG, "%s: sym_type=%d (F32, F64 or F16)", fn, sym_type);
    if (!modulation_ok(modulation)) return fail(TDB200_ERR_INVALID_ARG, "%s: modulation=%d (1, 2, 3, 4 or 6 bits per symbol)", fn, modulation);
    if (d->NL % modulation) return fail(TDB200_ERR_UNSUPPORTED, "%s: 3K+12 = %d is not a multiple of %d", fn, d->NL, modulation);
    return TDB200_OK;
}

int tdb200_decode_symbols_batch(tdb200_decoder *d, const void *sym_i, const void *sym_q, int sym_type, int mem, int n_cb,
                                int modulation, double kf, const tdb200_outputs *out, void *stream)
{
    if (!d || !sym_i || !sym_q || !out) return fail(TDB200_ERR_INVALID_ARG, "dec/sym/out is NULL");
    int s = check_sym("tdb200_decode_symbols_batch", d, sym_type, modulation);
    if (s) return s;
    if (!(kf > 0.0)) return fail(TDB200_ERR_INVALID_ARG, "kf must be positive");
    Source src;
    src.sym_i = sym_i; src.sym_q = sym_q; src.sym_type = sym_type; src.modulation = modulation; src.kf = kf;
    return decode_core(d, src, mem, n_cb, out, stream);
}

int tdb200_decode_rm_batch(tdb200_decoder *d, const void *e_llr, int llr_type, int mem, int n_cb, int E, int rv, int ncb,
                           const tdb200_outputs *out, void *stream)
{
    if (!d || !out || (!e_llr && n_cb > 0 && E > 0)) return fail(TDB200_ERR_INVALID_ARG, "dec/e_llr/out is NULL");
    if (llr_type < TDB200_LLR_F64 || llr_type > TDB200_LLR_F16) return fail(TDB200_ERR_INVALID_ARG, "llr_type=%d", llr_type);
    if (E < 0) return fail(TDB200_ERR_INVALID_ARG, "E=%d", E);
    Source src;
    src.llr = e_llr; src.llr_type = llr_type; src.rm_E = E;
    int s = rm_table(d, rv, ncb, &src.rm);
    if (s) return s;
    return decode_core(d, src, mem, n_cb, out, stream);
}

// A flat element-wise stage with host buffers: stage in, run, stage out (test-harness convenience;
// device callers are asynchronous on `stream`).
struct HostStage {
    std::vector<void *> dev;
    ~HostStage() { for (void *p : dev) cudaFree(p); }
    void *alloc(size_t bytes)
    {
        void *p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        dev.push_back(p);
        return p;
    }
};

int tdb200_modulate_flat(tdb200_decoder *d, const uint8_t *coded, void *sym_i, void *sym_q, int sym_type, int mem, size_t nb,
                         int modulation, void *stream)
{
    if (!d || !coded || !sym_i || !sym_q) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE) return fail(TDB200_ERR_INVALID_ARG, "mem=%d", mem);
    int s = check_sym("tdb200_modulate", d, sym_type, modulation);
    if (s) return s;
    if (nb % modulation) return fail(TDB200_ERR_INVALID_ARG, "%zu bits are not a whole number of %d-bit symbols", nb, modulation);
    if (nb == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    const size_t ns = nb / modulation, ssz = llr_elem_size(sym_type);
    if (mem == TDB200_MEM_DEVICE) {
        TDB_CUDA(launch_modulate(coded, sym_i, sym_q, sym_type, nb, modulation, st));
        return TDB200_OK;
    }
    HostStage hs;
    uint8_t *dc = static_cast<uint8_t *>(hs.alloc(nb));
    void *di = hs.alloc(ns * ssz), *dq = hs.alloc(ns * ssz);
    if (!dc || !di || !dq) return fail(TDB200_ERR_ALLOC, "device allocation failed");
    TDB_CUDA(cudaMemcpyAsync(dc, coded, nb, cudaMemcpyHostToDevice, st));
    TDB_CUDA(launch_modulate(dc, di, dq, sym_type, nb, modulation, st));
    TDB_CUDA(cudaMemcpyAsync(sym_i, di, ns * ssz, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaMemcpyAsync(sym_q, dq, ns * ssz, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaStreamSynchronize(st));
    return TDB200_OK;
}

int tdb200_modulate_batch(tdb200_decoder *d, const uint8_t *coded, void *sym_i, void *sym_q, int sym_type, int mem, int n_cb,
                          int modulation, void *stream)
{
    if (!d) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d", n_cb);
    return tdb200_modulate_flat(d, coded, sym_i, sym_q, sym_type, mem, (size_t)n_cb * d->NL, modulation, stream);
}

int tdb200_awgn_batch(tdb200_decoder *d, const void *x, void *y, int type, int mem, size_t n, double sigma, uint64_t seed, void *stream)
{
    if (!d || !x || !y) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE) return fail(TDB200_ERR_INVALID_ARG, "mem=%d", mem);
    if (type != TDB200_LLR_F32 && type != TDB200_LLR_F64 && type != TDB200_LLR_F16) return fail(TDB200_ERR_INVALID_ARG, "type=%d (F32, F64 or F16)", type);
    if (!(sigma >= 0.0)) return fail(TDB200_ERR_INVALID_ARG, "sigma must not be negative");
    if (n == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    if (mem == TDB200_MEM_DEVICE) {
        TDB_CUDA(launch_awgn(x, y, type, n, sigma, seed, st));
        return TDB200_OK;
    }
    HostStage hs;
    const size_t bytes = n * llr_elem_size(type);
    void *dx = hs.alloc(bytes);
    if (!dx) return fail(TDB200_ERR_ALLOC, "device allocation failed");
    TDB_CUDA(cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, st));
    TDB_CUDA(launch_awgn(dx, dx, type, n, sigma, seed, st));
    TDB_CUDA(cudaMemcpyAsync(y, dx, bytes, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaStreamSynchronize(st));
    return TDB200_OK;
}

int tdb200_demap_flat(tdb200_decoder *d, const void *sym_i, const void *sym_q, int sym_type, void *llr, int llr_type, int mem,
                      size_t n_llr, int modulation, double kf, void *stream)
{
    if (!d || !sym_i || !sym_q || !llr) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE) return fail(TDB200_ERR_INVALID_ARG, "mem=%d", mem);
    if (llr_type < TDB200_LLR_F64 || llr_type > TDB200_LLR_F16) return fail(TDB200_ERR_INVALID_ARG, "llr_type=%d", llr_type);
    int s = check_sym("tdb200_demap", d, sym_type, modulation);
    if (s) return s;
    if (!(kf > 0.0)) return fail(TDB200_ERR_INVALID_ARG, "kf must be positive");
    if (n_llr % modulation) return fail(TDB200_ERR_INVALID_ARG, "%zu soft bits are not a whole number of %d-bit symbols", n_llr, modulation);
    if (n_llr == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    DemapArgs a{};
    a.sym_type = sym_type; a.llr_type = llr_type; a.n_llr = n_llr; a.modulation = modulation; a.kf = kf;
    a.frac_bits = d->cfg.frac_bits ? d->cfg.frac_bits : 3;
    a.clip = std::min((1 << (a.frac_bits + 4)) - 1, 127);
    if (mem == TDB200_MEM_DEVICE) {
        a.sym_i = sym_i; a.sym_q = sym_q; a.llr = llr;
        TDB_CUDA(launch_demap(a, st));
        return TDB200_OK;
    }
    HostStage hs;
    const size_t sb = a.n_llr / modulation * llr_elem_size(sym_type), lb = a.n_llr * llr_elem_size(llr_type);
    void *di = hs.alloc(sb), *dq = hs.alloc(sb), *dl = hs.alloc(lb);
    if (!di || !dq || !dl) return fail(TDB200_ERR_ALLOC, "device allocation failed");
    TDB_CUDA(cudaMemcpyAsync(di, sym_i, sb, cudaMemcpyHostToDevice, st));
    TDB_CUDA(cudaMemcpyAsync(dq, sym_q, sb, cudaMemcpyHostToDevice, st));
    a.sym_i = di; a.sym_q = dq; a.llr = dl;
    TDB_CUDA(launch_demap(a, st));
    TDB_CUDA(cudaMemcpyAsync(llr, dl, lb, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaStreamSynchronize(st));
    return TDB200_OK;
}

int tdb200_demap_batch(tdb200_decoder *d, const void *sym_i, const void *sym_q, int sym_type, void *llr, int llr_type, int mem,
                       int n_cb, int modulation, double kf, void *stream)
{
    if (!d) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d", n_cb);
    return tdb200_demap_flat(d, sym_i, sym_q, sym_type, llr, llr_type, mem, (size_t)n_cb * d->NL, modulation, kf, stream);
}

int tdb200_rate_match_batch(tdb200_decoder *d, const uint8_t *coded, uint8_t *e_bits, int mem, int n_cb, int E, int rv, int ncb, void *stream)
{
    if (!d || !coded || (!e_bits && n_cb > 0 && E > 0)) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0 || E < 0 || (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE)) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d E=%d mem=%d", n_cb, E, mem);
    const tdb200_decoder::RmTable *t = nullptr;
    int s = rm_table(d, rv, ncb, &t);
    if (s) return s;
    if (n_cb == 0 || E == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    if (mem == TDB200_MEM_DEVICE) {
        TDB_CUDA(launch_rate_match(coded, e_bits, t->d_perm, t->nnn, d->NL, E, n_cb, st));
        return TDB200_OK;
    }
    HostStage hs;
    uint8_t *dc = static_cast<uint8_t *>(hs.alloc((size_t)n_cb * d->NL)), *de = static_cast<uint8_t *>(hs.alloc((size_t)n_cb * E));
    if (!dc || !de) return fail(TDB200_ERR_ALLOC, "device allocation failed");
    TDB_CUDA(cudaMemcpyAsync(dc, coded, (size_t)n_cb * d->NL, cudaMemcpyHostToDevice, st));
    TDB_CUDA(launch_rate_match(dc, de, t->d_perm, t->nnn, d->NL, E, n_cb, st));
    TDB_CUDA(cudaMemcpyAsync(e_bits, de, (size_t)n_cb * E, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaStreamSynchronize(st));
    return TDB200_OK;
}

int tdb200_rate_dematch_batch(tdb200_decoder *d, const void *e_llr, void *llr, int llr_type, int mem, int n_cb, int E, int rv, int ncb,
                              int accumulate, void *stream)
{
    if (!d || !llr || (!e_llr && n_cb > 0 && E > 0)) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0 || E < 0 || (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE)) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d E=%d mem=%d", n_cb, E, mem);
    if (llr_type < TDB200_LLR_F64 || llr_type > TDB200_LLR_F16) return fail(TDB200_ERR_INVALID_ARG, "llr_type=%d", llr_type);
    const tdb200_decoder::RmTable *t = nullptr;
    int s = rm_table(d, rv, ncb, &t);
    if (s) return s;
    if (n_cb == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    RmArgs a{};
    a.in_type = a.out_type = llr_type; a.inv = t->d_inv; a.nnn = t->nnn; a.NL = d->NL; a.E = E; a.n_cb = n_cb; a.accumulate = accumulate ? 1 : 0;
    a.frac_bits = d->cfg.frac_bits ? d->cfg.frac_bits : 3; a.clip = 127;
    if (mem == TDB200_MEM_DEVICE) {
        a.e_llr = e_llr; a.llr = llr;
        TDB_CUDA(launch_rate_dematch(a, st));
        return TDB200_OK;
    }
    HostStage hs;
    const size_t esz = llr_elem_size(llr_type), eb = (size_t)n_cb * E * esz, lb = (size_t)n_cb * d->NL * esz;
    void *de = hs.alloc(eb), *dl = hs.alloc(lb);
    if (!de || !dl) return fail(TDB200_ERR_ALLOC, "device allocation failed");
    TDB_CUDA(cudaMemcpyAsync(de, e_llr, eb, cudaMemcpyHostToDevice, st));
    if (accumulate) TDB_CUDA(cudaMemcpyAsync(dl, llr, lb, cudaMemcpyHostToDevice, st));
    a.e_llr = de; a.llr = dl;
    TDB_CUDA(launch_rate_dematch(a, st));
    TDB_CUDA(cudaMemcpyAsync(llr, dl, lb, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaStreamSynchronize(st));
    return TDB200_OK;
}

static int crc_common(tdb200_decoder *d, uint8_t *bits, int row_bits, int which, int attach, uint8_t *ok, int32_t *rem, int mem, int n_cb, void *stream)
{
    if (!d || !bits) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0 || (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE)) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d mem=%d", n_cb, mem);
    if (which != TDB200_CRC24A && which != TDB200_CRC24B) return fail(TDB200_ERR_INVALID_ARG, "which=%d (TDB200_CRC24A or TDB200_CRC24B)", which);
    const int K = row_bits ? row_bits : d->cfg.K;
    if (K <= 24 || K > (1 << 24)) return fail(TDB200_ERR_INVALID_ARG, "row_bits=%d: need 24 < row_bits <= 2^24", K);
    if (n_cb == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    CrcArgs a{};
    a.K = K; a.n_cb = n_cb; a.poly = (which == TDB200_CRC24A) ? 0x864CFBu : 0x800063u; a.attach = attach;
    if (mem == TDB200_MEM_DEVICE) {
        a.bits = bits; a.ok = ok; a.remainder = rem;
        TDB_CUDA(launch_crc24(a, st));
        return TDB200_OK;
    }
    HostStage hs;
    uint8_t *db = static_cast<uint8_t *>(hs.alloc((size_t)n_cb * K));
    uint8_t *dk = ok ? static_cast<uint8_t *>(hs.alloc(n_cb)) : nullptr;
    int32_t *dr = rem ? static_cast<int32_t *>(hs.alloc(sizeof(int32_t) * (size_t)n_cb)) : nullptr;
    if (!db || (ok && !dk) || (rem && !dr)) return fail(TDB200_ERR_ALLOC, "device allocation failed");
    TDB_CUDA(cudaMemcpyAsync(db, bits, (size_t)n_cb * K, cudaMemcpyHostToDevice, st));
    a.bits = db; a.ok = dk; a.remainder = dr;
    TDB_CUDA(launch_crc24(a, st));
    if (attach) TDB_CUDA(cudaMemcpyAsync(bits, db, (size_t)n_cb * K, cudaMemcpyDeviceToHost, st));
    if (ok) TDB_CUDA(cudaMemcpyAsync(ok, dk, n_cb, cudaMemcpyDeviceToHost, st));
    if (rem) TDB_CUDA(cudaMemcpyAsync(rem, dr, sizeof(int32_t) * (size_t)n_cb, cudaMemcpyDeviceToHost, st));
    TDB_CUDA(cudaStreamSynchronize(st));
    return TDB200_OK;
}

int tdb200_crc24_attach_batch(tdb200_decoder *d, uint8_t *bits, int row_bits, int which, int mem, int n_rows, void *stream)
{
    return crc_common(d, bits, row_bits, which, 1, nullptr, nullptr, mem, n_rows, stream);
}

int tdb200_crc24_check_batch(tdb200_decoder *d, const uint8_t *bits, int row_bits, int which, uint8_t *ok, int32_t *remainder, int mem,
                             int n_rows, void *stream)
{
    if (!ok && !remainder) return fail(TDB200_ERR_INVALID_ARG, "ok and remainder are both NULL");
    return crc_common(d, const_cast<uint8_t *>(bits), row_bits, which, 0, ok, remainder, mem, n_rows, stream);
}

int tdb200_segmentation(int B, tdb200_seg_info *info)
{
    if (!info) return fail(TDB200_ERR_INVALID_ARG, "info is NULL");
    if (B <= 0) return fail(TDB200_ERR_INVALID_ARG, "B=%d", B);
    const int Z = 6144;
    int L = 0, C = 1;
    long Bp = B;
    if (B > Z) { L = 24; C = (B + (Z - L) - 1) / (Z - L); Bp = (long)B + (long)C * L; }
    // K_plus: the smallest block size with C * K >= B'; K_minus: the next smaller one
    int Kp = 0, Km = 0;
    for (auto &r : kLte) {
        if ((long)C * r[0] >= Bp) { Kp = r[0]; break; }
        Km = r[0];
    }
    if (!Kp) return fail(TDB200_ERR_INVALID_ARG, "B=%d does not fit %d code blocks", B, C);
    info->C = C; info->K_plus = Kp; info->L = L;
    if (C == 1) { info->K_minus = 0; info->C_plus = 1; info->C_minus = 0; }
    else {
        info->K_minus = Km;
        info->C_minus = (int)(((long)C * Kp - Bp) / (Kp - Km));
        info->C_plus = C - info->C_minus;
    }
    info->F = (int)((long)info->C_plus * Kp + (long)info->C_minus * info->K_minus - Bp);
    return TDB200_OK;
}

int tdb200_encode_batch(tdb200_decoder *d, const uint8_t *bits, uint8_t *coded, int mem, int n_cb, void *stream)
{
    if (!d || !bits || !coded) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0 || (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE)) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d mem=%d", n_cb, mem);
    if (n_cb == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    const int K = d->cfg.K, NL = d->NL;
    EncodeArgs a{};
    a.pi = d->d_pi; a.K = K; a.n_cb = n_cb;
    if (mem == TDB200_MEM_DEVICE) {
        a.bits = bits; a.coded = coded;
        TDB_CUDA(launch_encode(a, st));
        return TDB200_OK;
    }
    uint8_t *db = nullptr, *dc = nullptr;
    TDB_CUDA(cudaMalloc(&db, (size_t)n_cb * K));
    if (cudaMalloc(&dc, (size_t)n_cb * NL) != cudaSuccess) { cudaFree(db); return fail(TDB200_ERR_ALLOC, "device allocation failed"); }
    cudaError_t e = cudaMemcpyAsync(db, bits, (size_t)n_cb * K, cudaMemcpyHostToDevice, st);
    a.bits = db; a.coded = dc;
    if (e == cudaSuccess) e = launch_encode(a, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(coded, dc, (size_t)n_cb * NL, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(db); cudaFree(dc);
    if (e != cudaSuccess) return fail(TDB200_ERR_CUDA, "tdb200_encode_batch: %s", cudaGetErrorString(e));
    return TDB200_OK;
}

int tdb200_channel_batch(tdb200_decoder *d, const uint8_t *coded, void *llr, int llr_type, int mem, int n_cb,
                         double sigma, uint64_t seed, void *stream)
{
    if (!d || !coded || !llr) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (n_cb < 0 || (mem != TDB200_MEM_HOST && mem != TDB200_MEM_DEVICE)) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d mem=%d", n_cb, mem);
    if (llr_type != TDB200_LLR_F32 && llr_type != TDB200_LLR_F64 && llr_type != TDB200_LLR_F16)
        return fail(TDB200_ERR_INVALID_ARG, "llr_type=%d (F32, F64 or F16)", llr_type);
    if (!(sigma > 0.0)) return fail(TDB200_ERR_INVALID_ARG, "sigma must be positive");
    if (n_cb == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    const size_t n = (size_t)n_cb * d->NL, esz = llr_elem_size(llr_type);
    ChannelArgs a{};
    a.n = n; a.sigma = (float)sigma; a.seed = seed;
    if (mem == TDB200_MEM_DEVICE) {
        a.coded = coded;
        TDB_CUDA(launch_channel(a, llr, llr_type, st));
        return TDB200_OK;
    }
    uint8_t *dc = nullptr;
    void *dl = nullptr;
    TDB_CUDA(cudaMalloc(&dc, n));
    if (cudaMalloc(&dl, n * esz) != cudaSuccess) { cudaFree(dc); return fail(TDB200_ERR_ALLOC, "device allocation failed"); }
    cudaError_t e = cudaMemcpyAsync(dc, coded, n, cudaMemcpyHostToDevice, st);
    a.coded = dc;
    if (e == cudaSuccess) e = launch_channel(a, dl, llr_type, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(llr, dl, n * esz, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dc); cudaFree(dl);
    if (e != cudaSuccess) return fail(TDB200_ERR_CUDA, "tdb200_channel_batch: %s", cudaGetErrorString(e));
    return TDB200_OK;
}

int tdb200_siso_batch(tdb200_decoder *d, const double *recs, const double *La, int terminated,
                      double *LLR, int mem, int n_cb, void *stream)
{
    if (!d || !recs || !La || !LLR) return fail(TDB200_ERR_INVALID_ARG, "NULL argument");
    if (d->cfg.algo != TDB200_ALGO_LOGMAP_F64) return fail(TDB200_ERR_UNSUPPORTED, "tdb200_siso_batch needs a TDB200_ALGO_LOGMAP_F64 decoder");
    if (n_cb < 0) return fail(TDB200_ERR_INVALID_ARG, "n_cb=%d", n_cb);
    d->launches_last = 0;
    if (n_cb == 0) return TDB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDB_DEVICE(d->cfg.device);
    const int T = d->T, chunk = d->cfg.max_batch;
    const bool host = (mem == TDB200_MEM_HOST);
    if (host) {
        int s = ensure(d->siso_in, d->siso_in_bytes, (size_t)chunk * 3 * T * 8);
        if (s) return s;
        if ((s = ensure_once(d->siso_out, 8 * (size_t)chunk * T))) return s;
    }
    for (int c0 = 0; c0 < n_cb; c0 += chunk) {
        const int n = std::min(chunk, n_cb - c0);
        Ref64SisoArgs a{};
        a.terminated = terminated; a.n_cb = n; a.T = T; a.ws = d->ws64;
        if (host) {
            double *din = static_cast<double *>(d->siso_in);
            TDB_CUDA(cudaMemcpyAsync(din, recs + (size_t)c0 * 2 * T, 8 * (size_t)n * 2 * T, cudaMemcpyHostToDevice, st));
            TDB_CUDA(cudaMemcpyAsync(din + (size_t)chunk * 2 * T, La + (size_t)c0 * T, 8 * (size_t)n * T, cudaMemcpyHostToDevice, st));
            a.recs = din; a.La = din + (size_t)chunk * 2 * T; a.LLR = static_cast<double *>(d->siso_out);
        } else {
            a.recs = recs + (size_t)c0 * 2 * T; a.La = La + (size_t)c0 * T; a.LLR = LLR + (size_t)c0 * T;
        }
        TDB_CUDA(launch_ref64_siso(a, st, &d->launches_last));
        if (host) {
            TDB_CUDA(cudaMemcpyAsync(LLR + (size_t)c0 * T, d->siso_out, 8 * (size_t)n * T, cudaMemcpyDeviceToHost, st));
            TDB_CUDA(cudaStreamSynchronize(st));
        }
    }
    return TDB200_OK;
}

}  // extern "C"
