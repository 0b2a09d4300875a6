// h2d_control.cpp -- the box's host-to-device ceiling: N concurrent pinned cudaMemcpyAsync H2D streams, one process per
// GPU (as torchrun launches bench.py), no kernel.  Control experiment for the end-to-end scaling of bench.py: if this
// aggregate stops growing with N, the end-to-end number through host buffers cannot grow either.
//
//   nvcc -O2 -o h2d_control h2d_control.cpp            (built by turbo_decoder_cuda_b200/build.py into lib/)
//   h2d_control [--gpus N] [--mb 302] [--reps 20] [--chunks 8] [--d2h-mb 25]
//
// Prints one JSON line: per-GPU and aggregate GB/s (aggregate = all bytes / the longest wall time).  --chunks splits each
// step's copy into that many cudaMemcpyAsync calls (the decoder's host path copies chunk by chunk); --d2h-mb adds a
// concurrent device-to-host stream of that size per step (the hard decisions going back).
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct Shared {
    pthread_barrier_t bar;
    double secs[16];
    double gbs[16];
    int err[16];
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int child(int g, Shared *sh, size_t bytes, int reps, int chunks, size_t d2h_bytes)
{
    int rc = 0;
    void *h = nullptr, *d = nullptr, *h2 = nullptr, *d2 = nullptr;
    cudaStream_t s = nullptr, s2 = nullptr;
    if (cudaSetDevice(g) != cudaSuccess) rc = 1;
    if (!rc && cudaMallocHost(&h, bytes) != cudaSuccess) rc = 2;
    if (!rc && cudaMalloc(&d, bytes) != cudaSuccess) rc = 3;
    if (!rc && d2h_bytes && (cudaMallocHost(&h2, d2h_bytes) != cudaSuccess || cudaMalloc(&d2, d2h_bytes) != cudaSuccess)) rc = 4;
    if (!rc) {
        memset(h, 1, bytes);
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    }
    const size_t chunk = (bytes / chunks) & ~(size_t)255;
    auto step = [&]() {
        for (int c = 0; c < chunks; c++)
            cudaMemcpyAsync((char *)d + c * chunk, (char *)h + c * chunk, chunk, cudaMemcpyHostToDevice, s);
        if (d2h_bytes) cudaMemcpyAsync(h2, d2, d2h_bytes, cudaMemcpyDeviceToHost, s2);
    };
    if (!rc) {
        for (int i = 0; i < 3; i++) step();
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(s2);
    }
    pthread_barrier_wait(&sh->bar);
    const double t0 = now();
    if (!rc) {
        for (int i = 0; i < reps; i++) step();
        if (cudaStreamSynchronize(s) != cudaSuccess) rc = 5;
        cudaStreamSynchronize(s2);
    }
    const double t1 = now();
    sh->secs[g] = t1 - t0;
    sh->gbs[g] = rc ? 0.0 : (double)(chunk * chunks) * reps / (t1 - t0) / 1e9;
    sh->err[g] = rc;
    pthread_barrier_wait(&sh->bar);
    return rc;
}

int main(int argc, char **argv)
{
    int gpus = 1, reps = 20, chunks = 8;
    double mb = 302.0, d2h_mb = 0.0;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--gpus")) gpus = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--mb")) mb = atof(argv[i + 1]);
        else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--chunks")) chunks = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--d2h-mb")) d2h_mb = atof(argv[i + 1]);
    }
    if (gpus < 1 || gpus > 16 || chunks < 1) return 2;
    Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) return 2;
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &at, gpus);
    const size_t bytes = (size_t)(mb * 1e6), d2h = (size_t)(d2h_mb * 1e6);
    for (int g = 0; g < gpus; g++) {  // fork BEFORE any CUDA call: one process (and one CUDA context) per GPU
        pid_t p = fork();
        if (p == 0) _exit(child(g, sh, bytes, reps, chunks, d2h));
    }
    int bad = 0;
    for (int g = 0; g < gpus; g++) {
        int st = 0;
        wait(&st);
        if (!WIFEXITED(st) || WEXITSTATUS(st)) bad++;
    }
    double worst = 0, sum = 0;
    for (int g = 0; g < gpus; g++) { worst = sh->secs[g] > worst ? sh->secs[g] : worst; sum += sh->gbs[g]; }
    const double total = (double)((bytes / chunks) & ~(size_t)255) * chunks * reps * gpus;
    printf("{\"control\": \"pinned cudaMemcpyAsync H2D, one process per GPU, no kernel\", \"n_gpus\": %d, \"mb_per_step\": %.1f, \"reps\": %d, "
           "\"chunks_per_step\": %d, \"d2h_mb_per_step\": %.1f, \"aggregate_gb_s\": %.2f, \"sum_of_per_gpu_gb_s\": %.2f, \"failed\": %d, \"per_gpu_gb_s\": [",
           gpus, mb, reps, chunks, d2h_mb, worst > 0 ? total / worst / 1e9 : 0.0, sum, bad);
    for (int g = 0; g < gpus; g++) printf("%s%.2f", g ? ", " : "", sh->gbs[g]);
    printf("]}\n");
    return bad ? 1 : 0;
}
