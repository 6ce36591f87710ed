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
// REGISTER FILE of the CTA's 512 threads (96 floats each) for the whole sequence; only the
// 128-float state vector moves through shared memory.  Two dependent mat-vecs per step
// (gates, then candidate on r*h), two __syncthreads per step.
#include "common.cuh"

namespace {

constexpr int H = 128;

__global__ void __launch_bounds__(512, 1)
bigru_kernel(const float* __restrict__ xp, const float* __restrict__ Wg_fw, const float* __restrict__ Wc_fw,
             const float* __restrict__ Wg_bw, const float* __restrict__ Wc_bw, float* __restrict__ out, int T) {
    const int b = blockIdx.x;
    const int dir = blockIdx.y;
    const int tid = threadIdx.x;
    const float* Wg = dir ? Wg_bw : Wg_fw;     // [128][256]  h-side rows of the gates kernel
    const float* Wc = dir ? Wc_bw : Wc_fw;     // [128][128]  (r*h)-side rows of the candidate kernel

    __shared__ __align__(16) float h_s[H];
    __shared__ __align__(16) float rh_s[H];
    __shared__ float u_s[H];

    // gates: column gcol (0..255), k half gk (64 k each); candidate: column ccol (0..127), k quarter ck (32 each)
    const int gcol = tid >> 1, gk = tid & 1;
    const int ccol = tid >> 2, ck = tid & 3;
    float wg[64], wc[32];
#pragma unroll
    for (int i = 0; i < 64; ++i) wg[i] = __ldg(Wg + (int64_t)(gk * 64 + i) * 256 + gcol);
#pragma unroll
    for (int i = 0; i < 32; ++i) wc[i] = __ldg(Wc + (int64_t)(ck * 32 + i) * 128 + ccol);

    if (tid < H) h_s[tid] = 0.0f;               // zero initial state (ops.py:112-115, s is None)
    __syncthreads();

    const int64_t seq = (int64_t)b * T;
    auto xrow = [&](int step) { int t = dir ? (T - 1 - step) : step; return xp + (seq + t) * 768 + dir * 384; };

    float xg_n = 0.f, xc_n = 0.f;
    if (T > 0) {
        const float* x0 = xrow(0);
        if (gk == 0) xg_n = __ldg(x0 + gcol);
        if (ck == 0) xc_n = __ldg(x0 + 256 + ccol);
    }
    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        const float xg = xg_n, xc = xc_n;
        if (step + 1 < T) {                     // prefetch next step's input products (independent of h)
            const float* xn = xrow(step + 1);
            if (gk == 0) xg_n = __ldg(xn + gcol);
            if (ck == 0) xc_n = __ldg(xn + 256 + ccol);
        }
        // ---- gates: h . Wg_h ----
        float a0 = 0.f, a1 = 0.f;
        const float4* h4 = reinterpret_cast<const float4*>(h_s + gk * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float4 hv = h4[i];
            a0 = fmaf(hv.x, wg[4 * i + 0], a0);
            a1 = fmaf(hv.y, wg[4 * i + 1], a1);
            a0 = fmaf(hv.z, wg[4 * i + 2], a0);
            a1 = fmaf(hv.w, wg[4 * i + 3], a1);
        }
        float acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (gk == 0) {
            float g = sigmoidf_acc(acc + xg);
            if (gcol < H) rh_s[gcol] = g * h_s[gcol];   // r * h
            else u_s[gcol - H] = g;                      // u
        }
        __syncthreads();
        // ---- candidate: (r*h) . Wc_h ----
        float c0 = 0.f, c1 = 0.f;
        const float4* r4 = reinterpret_cast<const float4*>(rh_s + ck * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 rv = r4[i];
            c0 = fmaf(rv.x, wc[4 * i + 0], c0);
            c1 = fmaf(rv.y, wc[4 * i + 1], c1);
            c0 = fmaf(rv.z, wc[4 * i + 2], c0);
            c1 = fmaf(rv.w, wc[4 * i + 3], c1);
        }
        float cacc = c0 + c1;
        cacc += __shfl_xor_sync(0xffffffffu, cacc, 1);
        cacc += __shfl_xor_sync(0xffffffffu, cacc, 2);
        if (ck == 0) {
            float c = tanhf_acc(cacc + xc);
            float u = u_s[ccol];
            float hn = u * h_s[ccol] + (1.0f - u) * c;
            h_s[ccol] = hn;
            out[(seq + t) * 256 + dir * H + ccol] = hn;
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
