// ittc_compat.cpp -- the reference's own entry points, re-exported on top of the C ABI.
//
// Defines, with the reference's exact C++ signatures (ITTC/main.h:13-24, so the mangled names are
// _Z13TurboDecodingPdPii and _Z15Log_MAP_decoderPdS_iS_i):
//     void TurboDecoding(double *flow_for_decode, int *flow_decoded, int flow_length);   log_map.cpp:1146
//     void Log_MAP_decoder(double *recs, double *La, int terminated, double *LLR, int len_total);  :898
//     void rate_match(int *input, int in_len, int *output, int out_len);                 main.h:23 (declared only)
//     void de_rate_match(double *input, double *output, int in_len, int out_len);        main.h:24 (declared only)
// and weak versions of TurboCodingInit / TurboCodingRelease / M_num_reg (log_map.cpp:349,1330,28) so
// that a build which still links the reference's log_map.cpp for the ENCODER side keeps those, and a
// build which drops log_map.cpp entirely still links.  See INTEGRATION.md for both recipes.
//
// Conventions reproduced from the reference:
//   * configuration comes from the caller-defined globals source_length / f1 / f2 (main.h:6-11);
//   * flow_for_decode is halved in place (log_map.cpp:1202-1205);
//   * flow_decoded receives N_ITERATION*K ints, iteration-major, natural order (:1264);
//   * failures print a message and exit(1) (the reference's malloc-failure behaviour, e.g. :909-913).
// N_ITERATION is a macro in the reference (log_map.h:30 -> 15); here it is TDB200_COMPAT_ITERS (default
// 15).  TDB200_COMPAT_ALGO selects the arithmetic: "logmap_f64" (default, reference-order fp64, the
// drop-in) or "maxlog_s16" (throughput mode; only the final decisions exist, every row gets them).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tdb200.h"

extern int source_length __attribute__((weak));
extern int f1 __attribute__((weak));
extern int f2 __attribute__((weak));

int M_num_reg __attribute__((weak)) = 3;

namespace {

struct State {
    tdb200_decoder *dec = nullptr;
    int K = 0, f1 = 0, f2 = 0, n_iter = 0, algo = 0;
    tdb200_decoder *siso = nullptr;
    int siso_T = 0;
} g;

[[noreturn]] void die(const char *what, int status)
{
    std::printf("\n tdb200: %s failed: %s (%s) \n", what, tdb200_status_string(status), tdb200_last_error());
    std::exit(1);
}

int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

void ensure_decoder(int K)
{
    const int n_iter = env_int("TDB200_COMPAT_ITERS", 15);
    const char *a = std::getenv("TDB200_COMPAT_ALGO");
    const int algo = (a && std::strcmp(a, "maxlog_s16") == 0) ? TDB200_ALGO_MAXLOG_S16
                     : ((a && std::strcmp(a, "logmap_s16") == 0) ? TDB200_ALGO_LOGMAP_S16 : TDB200_ALGO_LOGMAP_F64);
    const int qf1 = (&f1 && &source_length && source_length == K) ? f1 : 0;
    const int qf2 = (&f2 && &source_length && source_length == K) ? f2 : 0;
    if (g.dec && g.K == K && g.f1 == qf1 && g.f2 == qf2 && g.n_iter == n_iter && g.algo == algo) return;
    if (g.dec) tdb200_destroy(g.dec);
    g.dec = nullptr;
    tdb200_config cfg;
    tdb200_default_config(&cfg, K);
    cfg.f1 = qf1; cfg.f2 = qf2; cfg.n_iter = n_iter; cfg.algo = algo;
    cfg.device = env_int("TDB200_COMPAT_DEVICE", 0);
    cfg.max_batch = 4;
    int s = tdb200_create(&cfg, &g.dec);
    if (s != TDB200_OK) die("tdb200_create", s);
    g.K = K; g.f1 = qf1; g.f2 = qf2; g.n_iter = n_iter; g.algo = algo;
}

}  // namespace

void __attribute__((weak)) TurboCodingInit()
{
    if (&source_length) ensure_decoder(source_length);
}

void __attribute__((weak)) TurboCodingRelease()
{
    if (g.dec) tdb200_destroy(g.dec);
    if (g.siso) tdb200_destroy(g.siso);
    g = State();
}

void TurboDecoding(double *flow_for_decode, int *flow_decoded, int flow_length)
{
    const int K = (flow_length - 4 * 3) / 3;  // :1160
    ensure_decoder(K);
    tdb200_outputs out;
    std::memset(&out, 0, sizeof(out));
    // every decoder mode delivers the decisions after each iteration in the reference's flow_decoded layout (:1264)
    out.bits_iters = flow_decoded;
    int s = tdb200_decode_batch(g.dec, flow_for_decode, TDB200_LLR_F64, TDB200_MEM_HOST, 1, &out, nullptr);
    if (s != TDB200_OK) die("tdb200_decode_batch", s);
    for (int i = 0; i < flow_length; i++) flow_for_decode[i] *= 0.5;  // the reference's side effect, :1202-1205
}

void Log_MAP_decoder(double *recs_turbo, double *La_turbo, int terminated, double *LLR_all_turbo, int len_total)
{
    if (!g.siso || g.siso_T != len_total) {
        if (g.siso) tdb200_destroy(g.siso);
        g.siso = nullptr;
        tdb200_config cfg;
        tdb200_default_config(&cfg, len_total - 3);
        cfg.f1 = 1; cfg.f2 = 0;  // a SISO pass does not interleave; any permutation will do
        cfg.n_iter = 1; cfg.algo = TDB200_ALGO_LOGMAP_F64; cfg.max_batch = 4;
        cfg.device = env_int("TDB200_COMPAT_DEVICE", 0);
        int s = tdb200_create(&cfg, &g.siso);
        if (s != TDB200_OK) die("tdb200_create", s);
        g.siso_T = len_total;
    }
    int s = tdb200_siso_batch(g.siso, recs_turbo, La_turbo, terminated, LLR_all_turbo, TDB200_MEM_HOST, 1, nullptr);
    if (s != TDB200_OK) die("tdb200_siso_batch", s);
}

// The two stages the reference declares (ITTC/main.h:23-24) and calls from comments only
// (ITTC/main.cpp:196,204): with these definitions a maintainer can un-comment both call sites.
// TS 36.212 circular-buffer rate matching on the reference's multiplex order; redundancy version from
// TDB200_COMPAT_RV (default 0).  `in_len` / `out_len` are 3K+12 on the turbo-code side.
void rate_match(int *input, int in_len, int *output, int out_len)
{
    const int K = (in_len - 12) / 3;
    ensure_decoder(K);
    std::vector<uint8_t> c(in_len), e(out_len > 0 ? out_len : 1);
    for (int i = 0; i < in_len; i++) c[i] = (uint8_t)(input[i] & 1);
    int s = tdb200_rate_match_batch(g.dec, c.data(), e.data(), TDB200_MEM_HOST, 1, out_len, env_int("TDB200_COMPAT_RV", 0), 0, nullptr);
    if (s != TDB200_OK) die("tdb200_rate_match_batch", s);
    for (int i = 0; i < out_len; i++) output[i] = e[i];
}

void de_rate_match(double *input, double *output, int in_len, int out_len)
{
    const int K = (out_len - 12) / 3;
    ensure_decoder(K);
    int s = tdb200_rate_dematch_batch(g.dec, input, output, TDB200_LLR_F64, TDB200_MEM_HOST, 1, in_len, env_int("TDB200_COMPAT_RV", 0), 0, 0, nullptr);
    if (s != TDB200_OK) die("tdb200_rate_dematch_batch", s);
}
