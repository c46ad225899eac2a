// Which SM sub-partition (scheduler) does each warp of a small CTA land on?  %warpid is the hardware warp slot of the SM;
// slot % 4 is the sub-partition.  Launch shape = the reverse kernel's: 512 CTAs x 64 threads, 44 KB dynamic shared memory,
// 255 registers per thread (forced with a big live array).   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <map>
extern __shared__ double sm[];
__global__ void probe(int* out, int spin) {
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    sm[threadIdx.x] = threadIdx.x;
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    const int wpb = blockDim.x >> 5, w = blockIdx.x * wpb + (threadIdx.x >> 5);
    if ((threadIdx.x & 31) == 0) { out[w * 2] = smid; out[w * 2 + 1] = warpid; }
}
int main(int argc, char** argv) {
    int threads = argc > 1 ? atoi(argv[1]) : 64, blocks = argc > 2 ? atoi(argv[2]) : 512, regs_smem = argc > 3 ? atoi(argv[3]) : 44 * 1024;
    int* d; cudaMalloc(&d, blocks * (threads / 32) * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, regs_smem);
    probe<<<blocks, threads, regs_smem>>>(d, 2000000);
    cudaDeviceSynchronize();
    std::vector<int> h(blocks * (threads / 32) * 2);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    // per SM: list (block, warp-in-block, slot)
    std::map<int, std::vector<int>> per_sm;
    int wpb = threads / 32;
    for (int b = 0; b < blocks; b++) for (int w = 0; w < wpb; w++) { per_sm[h[(b * wpb + w) * 2]].push_back(w * 1000 + h[(b * wpb + w) * 2 + 1]); }
    int shown = 0;
    long mix[4][8] = {{0}};
    for (auto& kv : per_sm) {
        if (shown++ < 6) { printf("SM %3d:", kv.first); for (int v : kv.second) printf(" w%d@slot%d(sp%d)", v / 1000, v % 1000, (v % 1000) % 4); printf("\n"); }
        for (int v : kv.second) mix[(v % 1000) % 4][v / 1000]++;
    }
    for (int sp = 0; sp < 4; sp++) { printf("sub-partition %d: warps by index-in-CTA:", sp); for (int w = 0; w < wpb; w++) printf(" %ld", mix[sp][w]); printf("\n"); }
    return 0;
}
