// Which SM sub-partition (hardware warp slot %warpid mod 4) do the warps of co-resident CTAs get?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ubench_warpid tools/ubench_warpid.cu && ./ubench_warpid [threads] [smem_kb]
// Two CTAs per SM are forced with dynamic shared memory; every warp records (%smid, %warpid) and then
// spins long enough for the whole grid to be resident at once.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void probe(int *out, long long spin)
{
    extern __shared__ int sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) {
        int *o = out + 4 * (blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32);
        o[0] = smid; o[1] = warpid; o[2] = blockIdx.x; o[3] = threadIdx.x / 32;
    }
    const long long t0 = clock64();
    while (clock64() - t0 < spin) sm[threadIdx.x] = (int)spin;
}

int main(int argc, char **argv)
{
    const int threads = argc > 1 ? atoi(argv[1]) : 96, smem_kb = argc > 2 ? atoi(argv[2]) : 100;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = 2 * sms, wpc = threads / 32;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    int *d;
    cudaMalloc(&d, sizeof(int) * 4 * grid * wpc);
    probe<<<grid, threads, smem_kb * 1024>>>(d, 2000000);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    std::vector<int> h(4 * grid * wpc);
    cudaMemcpy(h.data(), d, sizeof(int) * h.size(), cudaMemcpyDeviceToHost);
    // histogram of the per-SM sub-partition loads
    std::vector<std::vector<int>> load(sms, std::vector<int>(4, 0));
    for (int i = 0; i < grid * wpc; i++) load[h[4 * i]][h[4 * i + 1] & 3]++;
    int hist[5][5][5][5] = {};
    for (int s = 0; s < sms; s++) hist[load[s][0] > 4 ? 4 : load[s][0]][load[s][1] > 4 ? 4 : load[s][1]][load[s][2] > 4 ? 4 : load[s][2]][load[s][3] > 4 ? 4 : load[s][3]]++;
    printf("threads per CTA %d, %d CTAs on %d SMs; per-SM warps on sub-partitions (0,1,2,3): count of SMs\n", threads, grid, sms);
    for (int a = 0; a < 5; a++) for (int b = 0; b < 5; b++) for (int c = 0; c < 5; c++) for (int e = 0; e < 5; e++)
        if (hist[a][b][c][e]) printf("  (%d,%d,%d,%d): %d\n", a, b, c, e, hist[a][b][c][e]);
    printf("first SM: ");
    for (int i = 0; i < grid * wpc; i++) if (h[4 * i] == h[0]) printf("[cta %d warp %d -> slot %d] ", h[4 * i + 2], h[4 * i + 3], h[4 * i + 1]);
    printf("\n");
    return 0;
}
