// probe_place.cu -- where do the warps of a 683 x 128-thread grid land?  Records %smid and
// %warpid (hardware warp slot; slot % 4 = scheduler) per (block, warp) while all blocks are resident.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <map>
#include <cuda_runtime.h>
__global__ void k(uint32_t *out, int spin) {
    uint32_t smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2] = smid; out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2 + 1] = wid; }
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
}
int main() {
    const int B = 683, W = 4;
    uint32_t *d; cudaMalloc(&d, B * W * 8);
    k<<<B, 32 * W>>>(d, 2000000);
    cudaDeviceSynchronize();
    std::vector<uint32_t> h(B * W * 2);
    cudaMemcpy(h.data(), d, B * W * 8, cudaMemcpyDeviceToHost);
    for (int b : {0, 1, 2, 3, 148, 149, 296, 444, 592, 682}) {
        printf("block %3d: sm %3u  warpids", b, h[b * W * 2]);
        for (int w = 0; w < W; ++w) printf(" %2u", h[(b * W + w) * 2 + 1]);
        printf("\n");
    }
    std::map<uint32_t, std::vector<int>> per_sm;
    for (int b = 0; b < B; ++b) per_sm[h[b * W * 2]].push_back(b);
    int shown = 0;
    for (auto &kv : per_sm) { if (shown++ >= 6) break; printf("sm %3u blocks:", kv.first); for (int b : kv.second) printf(" %d(w0 slot %u)", b, h[b * W * 2 + 1]); printf("\n"); }
    int hist[8] = {0};
    for (auto &kv : per_sm) hist[kv.second.size()]++;
    printf("blocks per SM histogram:"); for (int i = 0; i < 8; ++i) printf(" %d:%d", i, hist[i]); printf("\n");
    int mis = 0; for (int i = 0; i < B * W; ++i) if ((h[i * 2 + 1] & 3) != (uint32_t)(i % W)) mis++;
    printf("warps whose slot %% 4 != warp index: %d of %d\n", mis, B * W);
    return 0;
}
