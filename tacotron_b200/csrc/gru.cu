// gru.cu -- persistent bidirectional GRU (TF-1.2 GRUCell form), hidden size 128.
//
// Reference: tf.nn.bidirectional_dynamic_rnn(GRUCell(128), GRUCell(128), h) without
// sequence_length, models/ops.py:118-128; cell arithmetic SURVEY.md A.5:
//     [r,u] = sigmoid([x,h].Wg + bg)   c = tanh([x, r*h].Wc + bc)   h' = u*h + (1-u)*c
// (reset gate applied BEFORE the candidate product -- not the cuDNN form.)
//
// Design: the input-side products x.Wg[0:128], x.Wc[0:128] (+ biases) for all T steps are one
// big tensor-core GEMM done by the caller (xp, 768 columns = fw gates|fw cand|bw gates|bw cand).
// The serial part is one CTA per (utterance, direction): 64 CTAs at B=32, no inter-CTA
// communication at all.  The recurrent weights (128x256 + 128x128 fp32 = 192 KB) live in the
// REGISTER FILE of the CTA's 512 threads (96 floats each) for the whole sequence.
//
// Thread mapping (v2): shared-memory -> register bandwidth (128 B/clk/SM) was the bottleneck of
// the first version, where every thread re-read half of h each step (131 KB/step).  Now the 32
// lanes of a warp split K: lane l owns k = 4l..4l+3 (ONE 16-byte LDS of h per mat-vec), warp w
// owns 16 gate columns (8 candidate columns), and the per-column partial sums are combined with a
// halving butterfly (16 -> 8 -> 4 -> 2 -> 1 values over shfl_xor 16, 8, 4, 2, 1: 16 shuffles).
#include "common.cuh"

namespace {

constexpr int H = 128;

// d += a * b on two packed fp32 lanes (Blackwell FFMA2: one issue slot for two FMAs)
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

// Reduce N per-lane partial sums (N = 16 or 8) across the 32 lanes.  Each round halves the number of
// live values: a lane keeps the half selected by its lane bit and receives the partner's
// contribution for that half.  On return p[0] holds the full sum of column
//   N=16: c = 8*b4 + 4*b3 + 2*b2 + b1      N=8: c = 4*b4 + 2*b3 + b2      (b_i = bit i of lane)
template <int N>
__device__ __forceinline__ float butterfly(float (&p)[N], int lane) {
    int n = N;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            const bool hi = (lane & off) != 0;
            n >>= 1;
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                if (j < n) {
                    const float send = hi ? p[j] : p[j + n];
                    const float keep = hi ? p[j + n] : p[j];
                    p[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        } else {
            p[0] += __shfl_xor_sync(0xffffffffu, p[0], off);
        }
    }
    return p[0];
}

__global__ void __launch_bounds__(512, 1)
bigru_kernel(const float* __restrict__ xp, const float* __restrict__ Wg_fw, const float* __restrict__ Wc_fw,
             const float* __restrict__ Wg_bw, const float* __restrict__ Wc_bw, float* __restrict__ out, int T) {
    const int b = blockIdx.x;
    const int dir = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* Wg = dir ? Wg_bw : Wg_fw;     // [128][256]  h-side rows of the gates kernel
    const float* Wc = dir ? Wc_bw : Wc_fw;     // [128][128]  (r*h)-side rows of the candidate kernel

    __shared__ __align__(16) float h_s[H];
    __shared__ __align__(16) float rh_s[H];
    __shared__ float u_s[H];

    // weights: lane owns k = 4*lane + i; warp owns gate columns [16w,16w+16) and candidate columns [8w,8w+8)
    float wg[4][16], wc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int c = 0; c < 16; ++c) wg[i][c] = __ldg(Wg + (int64_t)(4 * lane + i) * 256 + warp * 16 + c);
#pragma unroll
        for (int c = 0; c < 8; ++c) wc[i][c] = __ldg(Wc + (int64_t)(4 * lane + i) * 128 + warp * 8 + c);
    }
    // after the butterflies: which column this lane finalises
    const int gcol = warp * 16 + (((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1));
    const bool g_owner = (lane & 1) == 0;
    const int ccol = warp * 8 + (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1));
    const bool c_owner = (lane & 3) == 0;

    if (tid < H) h_s[tid] = 0.0f;               // zero initial state (ops.py:112-115, s is None)
    __syncthreads();

    const int64_t seq = (int64_t)b * T;
    auto xrow = [&](int step) { int t = dir ? (T - 1 - step) : step; return xp + (seq + t) * 768 + dir * 384; };

    // input products are prefetched two steps ahead (they do not depend on h)
    float xg0 = 0.f, xc0 = 0.f, xg1 = 0.f, xc1 = 0.f;
    if (T > 0) { const float* x = xrow(0); if (g_owner) xg0 = __ldg(x + gcol); if (c_owner) xc0 = __ldg(x + 256 + ccol); }
    if (T > 1) { const float* x = xrow(1); if (g_owner) xg1 = __ldg(x + gcol); if (c_owner) xc1 = __ldg(x + 256 + ccol); }

    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        const float xg = xg0, xc = xc0;
        xg0 = xg1; xc0 = xc1;
        if (step + 2 < T) {
            const float* xn = xrow(step + 2);
            if (g_owner) xg1 = __ldg(xn + gcol);
            if (c_owner) xc1 = __ldg(xn + 256 + ccol);
        }
        // ---- gates: h . Wg_h ----
        {
            const float4 hv = *reinterpret_cast<const float4*>(h_s + 4 * lane);
            float2 p2[8];
            const float2 hx = make_float2(hv.x, hv.x), hy = make_float2(hv.y, hv.y), hz = make_float2(hv.z, hv.z), hw = make_float2(hv.w, hv.w);
#pragma unroll
            for (int c = 0; c < 8; ++c) p2[c] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) ffma2(p2[c], hx, make_float2(wg[0][2 * c], wg[0][2 * c + 1]));
#pragma unroll
            for (int c = 0; c < 8; ++c) ffma2(p2[c], hy, make_float2(wg[1][2 * c], wg[1][2 * c + 1]));
#pragma unroll
            for (int c = 0; c < 8; ++c) ffma2(p2[c], hz, make_float2(wg[2][2 * c], wg[2][2 * c + 1]));
#pragma unroll
            for (int c = 0; c < 8; ++c) ffma2(p2[c], hw, make_float2(wg[3][2 * c], wg[3][2 * c + 1]));
            float p[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) { p[2 * c] = p2[c].x; p[2 * c + 1] = p2[c].y; }
            const float acc = butterfly<16>(p, lane);
            if (g_owner) {
                const float g = sigmoidf_acc(acc + xg);
                if (gcol < H) rh_s[gcol] = g * h_s[gcol];   // r * h
                else u_s[gcol - H] = g;                      // u
            }
        }
        __syncthreads();
        // ---- candidate: (r*h) . Wc_h ----
        {
            const float4 rv = *reinterpret_cast<const float4*>(rh_s + 4 * lane);
            float2 p2[4];
            const float2 rx = make_float2(rv.x, rv.x), ry = make_float2(rv.y, rv.y), rz = make_float2(rv.z, rv.z), rw = make_float2(rv.w, rv.w);
#pragma unroll
            for (int c = 0; c < 4; ++c) p2[c] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) ffma2(p2[c], rx, make_float2(wc[0][2 * c], wc[0][2 * c + 1]));
#pragma unroll
            for (int c = 0; c < 4; ++c) ffma2(p2[c], ry, make_float2(wc[1][2 * c], wc[1][2 * c + 1]));
#pragma unroll
            for (int c = 0; c < 4; ++c) ffma2(p2[c], rz, make_float2(wc[2][2 * c], wc[2][2 * c + 1]));
#pragma unroll
            for (int c = 0; c < 4; ++c) ffma2(p2[c], rw, make_float2(wc[3][2 * c], wc[3][2 * c + 1]));
            float p[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) { p[2 * c] = p2[c].x; p[2 * c + 1] = p2[c].y; }
            const float cacc = butterfly<8>(p, lane);
            if (c_owner) {
                const float c = tanhf_acc(cacc + xc);
                const float u = u_s[ccol];
                const float hn = u * h_s[ccol] + (1.0f - u) * c;
                h_s[ccol] = hn;
                out[(seq + t) * 256 + dir * H + ccol] = hn;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int taco_bigru_fwd(const float* xp, const float* Wg_h_fw, const float* Wc_h_fw, const float* Wg_h_bw,
                              const float* Wc_h_bw, float* out, int B, int T, void* stream) {
    TACO_CHECK(xp && Wg_h_fw && Wc_h_fw && Wg_h_bw && Wc_h_bw && out, "taco_bigru_fwd: NULL pointer");
    TACO_CHECK(B >= 0 && T >= 0, "taco_bigru_fwd: negative size");
    if (B == 0 || T == 0) return 0;
    dim3 grid(B, 2);
    bigru_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(xp, Wg_h_fw, Wc_h_fw, Wg_h_bw, Wc_h_bw, out, T);
    TACO_LAUNCH_CHECK();
    return 0;
}
