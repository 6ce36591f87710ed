// mma.sync m16n8k8 tf32 latency (dependent chain) and issue throughput (independent accumulators) on B200
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void k(int iters, float* out, long long* cyc) {
    float d[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned a = threadIdx.x + 1, b = threadIdx.x * 3 + 7;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma(d[i], a, a + 1, a + 2, a + 3, b, b + 1);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    long long c;
    const int it = 2000;
    k<1><<<1, 32>>>(it, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 warp, dependent chain        : %.1f cycles / mma (latency)\n", (double)c / it);
    k<8><<<1, 32>>>(it, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 warp, 8 independent accs     : %.1f cycles / mma\n", (double)c / it / 8);
    k<8><<<1, 128>>>(it, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 warps (1/SMSP), 8 indep accs : %.1f cycles / mma per warp\n", (double)c / it / 8);
    k<8><<<1, 256>>>(it, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("8 warps (2/SMSP), 8 indep accs : %.1f cycles / mma per warp\n", (double)c / it / 8);
    k<3><<<1, 256>>>(it, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("8 warps, 3 indep accs          : %.1f cycles / mma per warp\n", (double)c / it / 3);
    return 0;
}
