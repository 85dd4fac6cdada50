// Where do the CTAs of a 2-CTA-cluster grid land?  Prints, per cluster, the SM ids of its two CTAs, for the persistent GEMM's launch
// shape (148 CTAs, 384 threads, ~225 KiB dynamic shared memory -> one CTA per SM).  Build: nvcc -arch=sm_100a (see tools/README.md).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(2, 1, 1) k(unsigned* out) {
    extern __shared__ unsigned char smem[];
    unsigned smid, nsmid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsmid));
    if (threadIdx.x == 0) { out[blockIdx.x] = smid; out[gridDim.x] = nsmid; smem[0] = 1; }
    // stay resident long enough that every CTA of the grid is placed at the same time
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < 2000000ull);
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* d;
    cudaMalloc(&d, (sms + 1) * 4);
    const int smem = 225 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<sms, 384, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned h[1024];
    cudaMemcpy(h, d, (sms + 1) * 4, cudaMemcpyDeviceToHost);
    printf("sms=%d nsmid=%u err=%s\ncluster: smid of CTA0, CTA1\n", sms, h[sms], cudaGetErrorString(e));
    for (int c = 0; c < sms / 2; ++c) printf("%3d: %3u %3u%s\n", c, h[2 * c], h[2 * c + 1], (h[2 * c] >> 1) == (h[2 * c + 1] >> 1) ? "" : "   <- not one TPC");
    return 0;
}
