// tdb200_burst.cpp -- a C++ caller of the C ABI (include/tdb200.h) with no Python and no torch in
// the process: one host thread per GPU, each with its own decoder handle and its own contiguous
// shard of the codeblocks (codeblocks are independent, nothing is exchanged between
// GPUs).  It plays the role of ITTC/main.cpp's Monte-Carlo loop (main.cpp:172-243) for a burst:
// random bits -> tdb200_encode_batch -> tdb200_channel_batch -> tdb200_decode_batch -> error count,
// everything device-resident, the decode timed with CUDA events.
//
//   tdb200_burst [--total N] [--gpus G] [--ebn0 dB] [--chunk C] [--early-term 0|1|2 (CRC24B)|3 (CRC24A)] [--K K] [--algo maxlog_s16|logmap_s16]
//                [--modulation 1|2|3|4|6] [--E bits-per-codeblock-after-rate-matching] [--rv 0..3]
//
// With --modulation > 1 or --E the loop is main.cpp's with the two commented-out stages restored:
// encode -> tdb200_rate_match_batch -> tdb200_modulate_flat -> tdb200_awgn_batch x2 -> tdb200_demap_flat ->
// tdb200_decode_rm_batch (timed: demap + de-rate-match + decode).
//
// Prints one JSON line.  Built by turbo_decoder_cuda_b200/build.py next to the libraries.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "tdb200.h"

namespace {

int g_algo = TDB200_ALGO_MAXLOG_S16;  // --algo maxlog_s16 | logmap_s16

struct Result {
    double ms = 0;
    long long bit_err = 0, frame_err = 0, iters = 0, n = 0;
    std::string error;
};

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) { r.error = std::string(#x) + ": " + cudaGetErrorString(e_); return; } \
    } while (0)
#define TK(x)                                                                                 \
    do {                                                                                      \
        int s_ = (x);                                                                         \
        if (s_ != TDB200_OK) { r.error = std::string(#x) + ": " + tdb200_last_error(); return; } \
    } while (0)

struct Chain { int modulation = 1, E = 0, rv = 0; };

void worker(int dev, int K, long long lo, long long hi, int chunk, double sigma, int early_term, Chain ch, Result &r)
{
    CK(cudaSetDevice(dev));
    tdb200_config cfg;
    tdb200_default_config(&cfg, K);
    cfg.device = dev; cfg.max_batch = chunk; cfg.early_term = early_term; cfg.algo = g_algo;
    tdb200_decoder *dec = nullptr;
    TK(tdb200_create(&cfg, &dec));
    const size_t NL = 3 * (size_t)K + 12;
    uint8_t *d_bits, *d_coded, *d_out;
    float *d_llr;
    int32_t *d_iters;
    CK(cudaMalloc(&d_bits, (size_t)chunk * K));
    CK(cudaMalloc(&d_coded, (size_t)chunk * NL));
    CK(cudaMalloc(&d_out, (size_t)chunk * K));
    CK(cudaMalloc(&d_llr, (size_t)chunk * NL * sizeof(float)));
    CK(cudaMalloc(&d_iters, (size_t)chunk * sizeof(int32_t)));
    // the optional rate-matching / mapping stages work on rows of E bits (E = 3K+12 without rate matching)
    const bool chain = ch.modulation > 1 || ch.E > 0;
    const size_t E = ch.E > 0 ? (size_t)ch.E : NL, NS = E / ch.modulation;
    uint8_t *d_tx = nullptr;
    float *d_si = nullptr, *d_sq = nullptr, *d_ri = nullptr, *d_rq = nullptr, *d_ellr = nullptr;
    if (chain) {
        CK(cudaMalloc(&d_tx, (size_t)chunk * E));
        CK(cudaMalloc(&d_si, (size_t)chunk * NS * sizeof(float)));
        CK(cudaMalloc(&d_sq, (size_t)chunk * NS * sizeof(float)));
        CK(cudaMalloc(&d_ri, (size_t)chunk * NS * sizeof(float)));
        CK(cudaMalloc(&d_rq, (size_t)chunk * NS * sizeof(float)));
        CK(cudaMalloc(&d_ellr, (size_t)chunk * E * sizeof(float)));
    }
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<uint8_t> h_bits((size_t)chunk * K), h_out((size_t)chunk * K);
    std::vector<int32_t> h_iters(chunk);
    std::mt19937_64 rng(0x9E3779B97F4A7C15ull ^ (unsigned long long)lo);
    {   // one untimed call on zeroed inputs: table uploads and scratch allocations happen here, not under the events
        const int n = (int)std::min<long long>(chunk, hi - lo);
        tdb200_outputs out;
        std::memset(&out, 0, sizeof(out));
        out.bits = d_out; out.iters_used = d_iters;
        if (!chain) {
            CK(cudaMemsetAsync(d_llr, 0, (size_t)n * NL * sizeof(float), st));
            TK(tdb200_decode_batch(dec, d_llr, TDB200_LLR_F32, TDB200_MEM_DEVICE, n, &out, st));
        } else {
            CK(cudaMemsetAsync(d_ri, 0, (size_t)n * NS * sizeof(float), st));
            CK(cudaMemsetAsync(d_rq, 0, (size_t)n * NS * sizeof(float), st));
            TK(tdb200_demap_flat(dec, d_ri, d_rq, TDB200_LLR_F32, d_ellr, TDB200_LLR_F32, TDB200_MEM_DEVICE, (size_t)n * E, ch.modulation, 1.0, st));
            TK(tdb200_decode_rm_batch(dec, d_ellr, TDB200_LLR_F32, TDB200_MEM_DEVICE, n, (int)E, ch.rv, 0, &out, st));
        }
        CK(cudaStreamSynchronize(st));
    }
    for (long long c0 = lo; c0 < hi; c0 += chunk) {
        const int n = (int)std::min<long long>(chunk, hi - c0);
        for (size_t i = 0; i < (size_t)n * K; i += 8) {  // 8 random bits per draw
            unsigned long long w = rng();
            for (int j = 0; j < 8 && i + j < (size_t)n * K; j++) h_bits[i + j] = (uint8_t)((w >> (8 * j)) & 1u);
        }
        CK(cudaMemcpyAsync(d_bits, h_bits.data(), (size_t)n * K, cudaMemcpyHostToDevice, st));
        if (early_term >= 2) {  // the CRC stopping rule needs blocks that end in a CRC: replace the last 24 bits
            TK(tdb200_crc24_attach_batch(dec, d_bits, 0, early_term == 2 ? TDB200_CRC24B : TDB200_CRC24A, TDB200_MEM_DEVICE, n, st));
            CK(cudaMemcpyAsync(h_bits.data(), d_bits, (size_t)n * K, cudaMemcpyDeviceToHost, st));
        }
        TK(tdb200_encode_batch(dec, d_bits, d_coded, TDB200_MEM_DEVICE, n, st));
        tdb200_outputs out;
        std::memset(&out, 0, sizeof(out));
        out.bits = d_out; out.iters_used = d_iters;
        if (!chain) {
            TK(tdb200_channel_batch(dec, d_coded, d_llr, TDB200_LLR_F32, TDB200_MEM_DEVICE, n, sigma, (uint64_t)(c0 + 1), st));
            CK(cudaEventRecord(e0, st));
            TK(tdb200_decode_batch(dec, d_llr, TDB200_LLR_F32, TDB200_MEM_DEVICE, n, &out, st));
            CK(cudaEventRecord(e1, st));
        } else {
            TK(tdb200_rate_match_batch(dec, d_coded, d_tx, TDB200_MEM_DEVICE, n, (int)E, ch.rv, 0, st));
            TK(tdb200_modulate_flat(dec, d_tx, d_si, d_sq, TDB200_LLR_F32, TDB200_MEM_DEVICE, (size_t)n * E, ch.modulation, st));
            TK(tdb200_awgn_batch(dec, d_si, d_ri, TDB200_LLR_F32, TDB200_MEM_DEVICE, (size_t)n * NS, sigma, (uint64_t)(2 * c0 + 1), st));
            TK(tdb200_awgn_batch(dec, d_sq, d_rq, TDB200_LLR_F32, TDB200_MEM_DEVICE, (size_t)n * NS, sigma, (uint64_t)(2 * c0 + 2), st));
            CK(cudaEventRecord(e0, st));
            TK(tdb200_demap_flat(dec, d_ri, d_rq, TDB200_LLR_F32, d_ellr, TDB200_LLR_F32, TDB200_MEM_DEVICE, (size_t)n * E, ch.modulation,
                                 1.0 / (2.0 * sigma * sigma), st));
            TK(tdb200_decode_rm_batch(dec, d_ellr, TDB200_LLR_F32, TDB200_MEM_DEVICE, n, (int)E, ch.rv, 0, &out, st));
            CK(cudaEventRecord(e1, st));
        }
        CK(cudaMemcpyAsync(h_out.data(), d_out, (size_t)n * K, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_iters.data(), d_iters, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        r.ms += ms;
        for (int c = 0; c < n; c++) {  // the reference's error counting, ITTC/main.cpp:224-237
            long long e = 0;
            for (int i = 0; i < K; i++) e += h_out[(size_t)c * K + i] != h_bits[(size_t)c * K + i];
            r.bit_err += e;
            r.frame_err += e > 0;
            r.iters += h_iters[c];
        }
        r.n += n;
    }
    cudaFree(d_bits); cudaFree(d_coded); cudaFree(d_out); cudaFree(d_llr); cudaFree(d_iters);
    cudaFree(d_tx); cudaFree(d_si); cudaFree(d_sq); cudaFree(d_ri); cudaFree(d_rq); cudaFree(d_ellr);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
    tdb200_destroy(dec);
}

}  // namespace

int main(int argc, char **argv)
{
    long long total = 65536;
    int gpus = 0, K = 6144, chunk = 8192, early_term = 0;
    double ebn0 = 1.0;
    Chain ch;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        if (a == "--total") total = std::atoll(argv[i + 1]);
        else if (a == "--gpus") gpus = std::atoi(argv[i + 1]);
        else if (a == "--K") K = std::atoi(argv[i + 1]);
        else if (a == "--chunk") chunk = std::atoi(argv[i + 1]);
        else if (a == "--ebn0") ebn0 = std::atof(argv[i + 1]);
        else if (a == "--early-term") early_term = std::atoi(argv[i + 1]);
        else if (a == "--modulation") ch.modulation = std::atoi(argv[i + 1]);
        else if (a == "--E") ch.E = std::atoi(argv[i + 1]);
        else if (a == "--rv") ch.rv = std::atoi(argv[i + 1]);
        else if (a == "--algo") g_algo = (std::string(argv[i + 1]) == "logmap_s16") ? TDB200_ALGO_LOGMAP_S16 : TDB200_ALGO_MAXLOG_S16;
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        std::fprintf(stderr, "tdb200_burst: no CUDA device (there is no CPU path)\n");
        return 1;
    }
    if (gpus <= 0 || gpus > ndev) gpus = ndev;
    if (ch.modulation != 1 && ch.modulation != 2 && ch.modulation != 3 && ch.modulation != 4 && ch.modulation != 6) {
        std::fprintf(stderr, "tdb200_burst: --modulation must be 1, 2, 3, 4 or 6\n");
        return 2;
    }
    if ((ch.E > 0 ? ch.E : 3 * K + 12) % ch.modulation) { std::fprintf(stderr, "tdb200_burst: E must be a multiple of the modulation order\n"); return 2; }
    // sigma at the code rate K/E and MODULATION bits per symbol, ITTC/main.cpp:47,174
    const double rate = K / (ch.E > 0 ? (double)ch.E : 3.0 * K + 12.0);
    const double sigma = std::pow(10.0, -ebn0 / 20.0) * std::sqrt(0.5 / (rate * ch.modulation));
    std::vector<Result> res(gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; g++) {
        const long long base = total / gpus, rem = total % gpus;
        const long long lo = g * base + std::min<long long>(g, rem), hi = lo + base + (g < rem ? 1 : 0);
        th.emplace_back(worker, g, K, lo, hi, chunk, sigma, early_term, ch, std::ref(res[g]));
    }
    for (auto &t : th) t.join();
    double ms_max = 0;
    long long be = 0, fe = 0, it = 0, n = 0;
    for (auto &r : res) {
        if (!r.error.empty()) { std::fprintf(stderr, "tdb200_burst: %s\n", r.error.c_str()); return 1; }
        ms_max = std::max(ms_max, r.ms);
        be += r.bit_err; fe += r.frame_err; it += r.iters; n += r.n;
    }
    std::printf("{\"tool\": \"tdb200_burst (C++ over the C ABI, one host thread per GPU)\", \"algo\": \"%s\", \"n_gpus\": %d, \"K\": %d, "
                "\"modulation\": %d, \"E\": %d, \"rv\": %d, \"codeblocks\": %lld, \"ebn0_db\": %.2f, \"early_term\": %d, \"decode_ms_max_over_gpus\": %.3f, "
                "\"gbit_s\": %.3f, \"bit_errors\": %lld, \"frame_errors\": %lld, \"mean_iters\": %.3f}\n",
                g_algo == TDB200_ALGO_LOGMAP_S16 ? "logmap_s16" : "maxlog_s16", gpus, K, ch.modulation, ch.E > 0 ? ch.E : 3 * K + 12, ch.rv, n, ebn0, early_term, ms_max, (double)n * K / (ms_max * 1e-3) / 1e9, be, fe, n ? (double)it / n : 0.0);
    return 0;
}
