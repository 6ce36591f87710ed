// Micro-benchmarks that size the decoder's exchange protocol on B200:
//   A: dependent-load latency through L2 (ld.relaxed.gpu vs ld.global.cg)
//   B: store->poll ping-pong between two CTAs through L2 (one hop = half a round trip)
//   C: the same ping-pong inside a 2-CTA cluster through distributed shared memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void chase(const uint64_t* buf, int n, int mode, uint64_t* out, long long* cyc) {
    uint64_t idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        uint64_t v;
        if (mode == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(buf + idx));
        else if (mode == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(buf + idx));
        else asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(buf + idx));
        idx = v;
    }
    long long t1 = clock64();
    out[0] = idx; cyc[0] = t1 - t0;
}

// ping-pong: CTA `a` and CTA `b` (other CTAs exit); flags are 64-bit LL words
__global__ void pingpong(uint64_t* fa, uint64_t* fb, int a, int b, int iters, long long* cyc) {
    if (threadIdx.x != 0) return;
    if ((int)blockIdx.x == a) {
        long long t0 = clock64();
        for (uint64_t i = 1; i <= (uint64_t)iters; ++i) {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(fa), "l"(i) : "memory");
            uint64_t v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(fb) : "memory"); } while (v != i);
        }
        cyc[0] = clock64() - t0;
    } else if ((int)blockIdx.x == b) {
        for (uint64_t i = 1; i <= (uint64_t)iters; ++i) {
            uint64_t v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(fa) : "memory"); } while (v != i);
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(fb), "l"(i) : "memory");
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int iters, long long* cyc) {
    __shared__ uint64_t flag;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    if (threadIdx.x == 0) flag = 0;
    cl.sync();
    volatile uint64_t* mine = &flag;
    uint64_t* peer = cl.map_shared_rank(&flag, rank ^ 1);
    if (threadIdx.x == 0) {
        if (rank == 0) {
            long long t0 = clock64();
            for (uint64_t i = 1; i <= (uint64_t)iters; ++i) {
                *(volatile uint64_t*)peer = i;
                while (*mine != i) {}
            }
            cyc[0] = clock64() - t0;
        } else {
            for (uint64_t i = 1; i <= (uint64_t)iters; ++i) {
                while (*mine != i) {}
                *(volatile uint64_t*)peer = i;
            }
        }
    }
    cl.sync();
}

int main() {
    const int N = 1 << 19;              // 4 MB of u64: L2 resident
    uint64_t* h = (uint64_t*)malloc(N * 8);
    // random cyclic permutation with a 128-byte granularity
    const int lines = N / 16;
    int* perm = (int*)malloc(lines * sizeof(int));
    for (int i = 0; i < lines; ++i) perm[i] = i;
    srand(1);
    for (int i = lines - 1; i > 0; --i) { int j = rand() % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (int i = 0; i < N; ++i) h[i] = 0;
    for (int i = 0; i < lines; ++i) h[(size_t)perm[i] * 16] = (uint64_t)perm[(i + 1) % lines] * 16;
    uint64_t *d, *out; long long* cyc;
    cudaMalloc(&d, N * 8); cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    cudaMemcpy(d, h, N * 8, cudaMemcpyHostToDevice);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock attr %d kHz\n", clk);
    const char* names[3] = {"ld.relaxed.gpu", "ld.global.cg", "ld.volatile"};
    for (int mode = 0; mode < 3; ++mode) {
        chase<<<1, 1>>>(d, 2000, mode, out, cyc);          // warm
        chase<<<1, 1>>>(d, 20000, mode, out, cyc);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("A dependent load %-16s %.1f cycles/load\n", names[mode], (double)c / 20000);
    }
    uint64_t* flags; cudaMalloc(&flags, 4096); cudaMemset(flags, 0, 4096);
    int pairs[4][2] = {{0, 1}, {0, 2}, {0, 73}, {0, 147}};
    for (int p = 0; p < 4; ++p) {
        cudaMemset(flags, 0, 4096);
        pingpong<<<148, 32>>>(flags, flags + 64, pairs[p][0], pairs[p][1], 2000, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("B L2 ping-pong CTA %d <-> %d: %.1f cycles/round trip (%.1f per hop) %s\n", pairs[p][0], pairs[p][1], (double)c / 2000,
               (double)c / 4000, cudaGetErrorString(e));
    }
    pingpong_dsmem<<<2, 32>>>(2000, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("C DSMEM ping-pong (cluster of 2): %.1f cycles/round trip (%.1f per hop) %s\n", (double)c / 2000, (double)c / 4000, cudaGetErrorString(e));
    return 0;
}
