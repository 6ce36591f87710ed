// decoder.cu -- the whole autoregressive attention decoder as ONE persistent cooperative kernel.
//
// Reference: models/tacotron.py:46-105 (create_decoder) + :136-138 (dynamic_decode) over the
// TF-1.2 contrib.seq2seq / contrib.rnn classes.  Per decoder step t:
//   P1  p1   = relu(x_last . W1 + b1) (*keep1)            x_last = last 80 of the 80r-wide input
//   P2  p2   = relu(p1 . W2 + b2) (*keep2)                 Tacotron.pre_net, tacotron.py:38-44,64-71
//   IN  z    = [p2, attn] . W_in + b_in                    InputProjectionWrapper
//   G_i [r,u]= sigmoid([x_i, h_i] . Wg_i + bg_i)           GRUCell x3 (MultiRNNCell), x_1 = z, x_i = h_{i-1}
//   C_i c    = tanh([x_i, r*h_i] . Wc_i + bc_i); h_i = u*h_i + (1-u)*c
//       s    = z + h_3                                     ResidualWrapper (around the 3-stack)
//   OUT y_t  = s . W_out + b_out                           OutputProjectionWrapper -> seq2seq_output
//   Q   q    = y_t . W_q                                   BahdanauAttention query layer (query = cell OUTPUT)
//   ATT e_j  = sum_d v_d tanh(keys_jd + q_d), j >= text_length masked; softmax; ctx = a . values
//   AL  attn = [y_t, ctx] . W_a                            AttentionWrapper attention_layer (no bias)
//   next input: InferenceHelper -> y_t ; TrainingHelper -> mel[:, t+1] ; ScheduledOutput -> per-row mix
//
// Design (B <= 32 utterances per launch):
//   * grid = 128 CTAs x 256 threads, launched cooperatively, one CTA per SM, alive for all T steps.
//   * every dense stage [32 x K] . [K x N] is split 4 row-groups x 32 column-slices: CTA (rg, cs)
//     owns 8 rows and N/32 columns.  Its K x NC weight slice is contiguous in the packed buffer
//     (taco_decoder_pack) and is either RESIDENT in shared memory for the whole kernel or streamed
//     from L2 by a bulk async copy (cp.async.bulk + mbarrier) one stage ahead, double buffered.
//   * the 13 dependent stages of a step are separated by a software grid barrier (release/acquire
//     counter); stage outputs are exchanged through L2 (ld.global.cg).
//   * attention: CTA (utterance, quarter of Tx) keeps its keys/values slice in shared memory for
//     all steps, scores+partial softmax+partial context per quarter, flash-style merge of the four
//     quarters by the consumer stage.
//   * %globaltimer stamp per step for the decoder-step latency metric.
#include <cooperative_groups.h>
#include "common.cuh"

namespace {

constexpr int NCTA = 128;
constexpr int NTHR = 256;
constexpr int RG = 4;          // row groups
constexpr int RPG = 8;         // rows per group
constexpr int NS = 32;         // column slices
constexpr int BPAD = RG * RPG; // 32
constexpr int U = 256;         // decoder units
constexpr int AU = 256;        // attention units
constexpr int ENC = 256;       // memory depth
constexpr int MF = 80;
constexpr int NSTAGE = 13;
constexpr int ACT_LD = 660;    // 656 + 4 (row stride of the staged activations; 660 % 32 == 20 -> conflict-free LDS.128)
constexpr int KV_LD = 260;     // padded row stride for keys in smem
constexpr int STREAM_FLOATS = 512 * 16;   // largest weight slice (K=512, NC=16)

enum StageKind { K_P1 = 0, K_P2, K_IN, K_G1, K_C1, K_G2, K_C2, K_G3, K_C3, K_OUT, K_Q, K_ATT, K_AL };

struct StageDesc {
    int K0, K1;        // K = K0 + K1 (two concatenated sources)
    int N;             // true output width
    int NC;            // columns per slice (4, 8 or 16); padded width = NC*32
    int64_t w_off;     // float offset of slice 0 in the packed buffer
    int res_off;       // float offset inside the resident smem region, or -1 = streamed
};

struct DecLayout {     // workspace float offsets (all [32][ld] row-major)
    int64_t p1, p2, attn, z, h[3], rh, u, s, ybuf, q, att_ms, att_ctx, ctr;
    int64_t total;
};

struct DecParams {
    StageDesc st[NSTAGE];
    DecLayout ws;
    taco_decoder_args a;
    // biases / small vectors straight from the TF-layout tensors
    const float *pre_b1, *pre_b2, *in_b, *gru_bg[3], *gru_bc[3], *out_b, *att_v;
    int OUT;           // 80*r
    int Tq;            // Tx/4
    int res_floats;    // size of resident region
    int smem_kv_off, smem_res_off, smem_stream_off, smem_total_floats;
};

__host__ __device__ inline void stage_dims(int kind, int OUT, int& K0, int& K1, int& N, int& NC) {
    switch (kind) {
        case K_P1: K0 = MF;  K1 = 0;   N = 256; NC = 8;  break;
        case K_P2: K0 = 256; K1 = 0;   N = 128; NC = 4;  break;
        case K_IN: K0 = 128; K1 = AU;  N = U;   NC = 8;  break;
        case K_G1: case K_G2: case K_G3: K0 = U; K1 = U; N = 2 * U; NC = 16; break;
        case K_C1: case K_C2: case K_C3: K0 = U; K1 = U; N = U;     NC = 8;  break;
        case K_OUT: K0 = U;  K1 = 0;   N = OUT; NC = (OUT + NS - 1) / NS; NC = (NC <= 4) ? 4 : (NC <= 8) ? 8 : 16; break;
        case K_Q:  K0 = OUT; K1 = 0;   N = AU;  NC = 8;  break;
        case K_AL: K0 = OUT; K1 = ENC; N = AU;  NC = 8;  break;
        default:   K0 = K1 = N = 0; NC = 4; break;      // K_ATT has no weight slice
    }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        red_release_add(ctr, 1u);
        unsigned int spins = 0;
        while (ld_acquire(ctr) < target) {
            if (++spins > (1u << 26)) __trap();      // a lost CTA must not hang the GPU
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(NTHR, 1) decoder_kernel(const DecParams P) {
    extern __shared__ __align__(16) float smem[];
    // smem map (floats): [0,64) barriers | act_s 8*ACT_LD | red_s 8*32*4 | small 1024 | kv | resident | stream 2x
    uint64_t* wbar = reinterpret_cast<uint64_t*>(smem);          // [0],[1] stream buffers, [2] resident/kv preload
    float* act_s = smem + 64;
    float* red_s = act_s + RPG * ACT_LD;
    float* small_s = red_s + 8 * 32 * 4;                          // q_s[256] | e_s[<=64] | p_s[<=64] | misc
    float* keys_s = smem + P.smem_kv_off;
    float* vals_s = keys_s + P.Tq * KV_LD;
    float* res_s = smem + P.smem_res_off;
    float* stream_s = smem + P.smem_stream_off;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x;
    const int rg = cta & 3, cs = cta >> 2;             // dense stages
    const int arow = cta >> 2, aq = cta & 3;           // attention stage: utterance, quarter
    const taco_decoder_args& A = P.a;
    float* ws = reinterpret_cast<float*>(A.workspace);
    unsigned int* ctr = reinterpret_cast<unsigned int*>(ws + P.ws.ctr);
    const int B = A.B, T = A.T, OUT = P.OUT, Tq = P.Tq, Tx = A.Tx;

    if (tid == 0) {
        mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); mbar_init(&wbar[2], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- one-time preload: resident weight slices (bulk copies) ----
    if (tid == 0) {
        uint32_t bytes = 0;
        for (int s = 0; s < NSTAGE; ++s) {
            const StageDesc& d = P.st[s];
            if (s == K_ATT || d.res_off < 0) continue;
            uint32_t nb = (uint32_t)((d.K0 + d.K1) * d.NC * 4);
            bytes += nb;
        }
        if (bytes) {
            mbar_arrive_expect_tx(&wbar[2], bytes);
            for (int s = 0; s < NSTAGE; ++s) {
                const StageDesc& d = P.st[s];
                if (s == K_ATT || d.res_off < 0) continue;
                uint32_t nb = (uint32_t)((d.K0 + d.K1) * d.NC * 4);
                bulk_load(res_s + d.res_off, A.packed + d.w_off + (int64_t)cs * (d.K0 + d.K1) * d.NC, nb, &wbar[2]);
            }
        } else {
            mbar_arrive(&wbar[2]);
        }
    }
    // keys / values slice of (arow, aq) -> smem (padded rows), zero for utterances >= B
    for (int i = tid; i < Tq * (ENC / 4); i += NTHR) {
        int j = i / (ENC / 4), d4 = (i % (ENC / 4)) * 4;
        float4 kk = make_float4(0, 0, 0, 0), vv = kk;
        if (arow < B) {
            int64_t g = ((int64_t)arow * Tx + aq * Tq + j) * ENC + d4;
            kk = __ldg(reinterpret_cast<const float4*>(A.keys + g));
            vv = __ldg(reinterpret_cast<const float4*>(A.values + g));
        }
        *reinterpret_cast<float4*>(keys_s + j * KV_LD + d4) = kk;
        *reinterpret_cast<float4*>(vals_s + j * ENC + d4) = vv;
    }
    const int my_len = (arow < B) ? A.text_length[arow] : 0;
    mbar_wait(&wbar[2], 0);
    __syncthreads();

    // streamed-slice bookkeeping: issue order == consume order; buffer = (index) & 1.  At most two slices
    // are in flight (the one the current stage will consume and the one after it).  A buffer is re-filled
    // only at the top of an iteration, when every thread has passed the post-compute __syncthreads of the
    // iteration that last read it.
    uint32_t n_issued = 0, n_consumed = 0;
    const int total_iters = T * NSTAGE;
    int next_issue_it = 0;
    auto is_streamed = [&](int s) { return s != K_ATT && P.st[s].res_off < 0; };
    auto pump = [&]() {
        while (n_issued - n_consumed < 2) {
            while (next_issue_it < total_iters && !is_streamed(next_issue_it % NSTAGE)) ++next_issue_it;
            if (next_issue_it >= total_iters) break;
            if (tid == 0) {
                const StageDesc& d = P.st[next_issue_it % NSTAGE];
                const uint32_t nb = (uint32_t)((d.K0 + d.K1) * d.NC * 4);
                const int buf = n_issued & 1;
                fence_proxy_async();
                mbar_arrive_expect_tx(&wbar[buf], nb);
                bulk_load(stream_s + buf * STREAM_FLOATS, A.packed + d.w_off + (int64_t)cs * (d.K0 + d.K1) * d.NC, nb, &wbar[buf]);
            }
            ++n_issued; ++next_issue_it;
        }
    };

    unsigned int bar_target = 0;
    for (int it = 0; it < total_iters; ++it) {
        const int t = it / NSTAGE;
        const int s = it - t * NSTAGE;
        pump();

        // ---- wait until the producers of this stage's inputs are done (grid-wide) ----
        if (it > 0) { bar_target += NCTA; grid_barrier(ctr, bar_target); }
        if (s == 0 && cta == 0 && tid == 0 && A.step_ns) A.step_ns[t] = globaltimer_ns();

        if (s == K_ATT) {
            // =============== attention scores / partial softmax / partial context ===============
            float* q_s = small_s;             // [256]
            float* v_s = small_s + 256;       // [256]
            float* e_s = small_s + 512;       // [Tq]
            float* p_s = small_s + 512 + 64;  // [Tq]
            float* red2 = small_s + 512 + 128; // [8]
            if (arow < B) {
                for (int i = tid; i < AU; i += NTHR) { q_s[i] = __ldcg(ws + P.ws.q + (int64_t)arow * AU + i); v_s[i] = __ldg(P.att_v + i); }
            }
            __syncthreads();
            if (arow < B) {
                const int dsl = tid & 7;
                // warp-uniform trip count (each warp covers 4 consecutive j per pass): the full-mask
                // shuffles below must be executed by all 32 lanes even when Tq is not a multiple of 4
                for (int j0 = warp * 4; j0 < Tq; j0 += NTHR / 8) {
                    const int j = j0 + (lane >> 3);
                    const bool valid = j < Tq;
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!valid) break;
                        const int d0 = i * 32 + dsl * 4;
                        float4 kk = *reinterpret_cast<const float4*>(keys_s + j * KV_LD + d0);
                        float4 qq = *reinterpret_cast<const float4*>(q_s + d0);
                        float4 vv = *reinterpret_cast<const float4*>(v_s + d0);
                        acc = fmaf(vv.x, tanhf_acc(kk.x + qq.x), acc);
                        acc = fmaf(vv.y, tanhf_acc(kk.y + qq.y), acc);
                        acc = fmaf(vv.z, tanhf_acc(kk.z + qq.z), acc);
                        acc = fmaf(vv.w, tanhf_acc(kk.w + qq.w), acc);
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    if (valid && dsl == 0) e_s[j] = (aq * Tq + j < my_len) ? acc : -INFINITY;
                }
            }
            __syncthreads();
            if (arow < B && warp == 0) {
                float m = -INFINITY;
                for (int j = lane; j < Tq; j += 32) m = fmaxf(m, e_s[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float ssum = 0.f;
                for (int j = lane; j < Tq; j += 32) {
                    float p = (m == -INFINITY) ? 0.f : __expf(e_s[j] - m);
                    p_s[j] = p;
                    ssum += p;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                if (lane == 0) {
                    float* ms = ws + P.ws.att_ms + ((int64_t)arow * 4 + aq) * 2;
                    ms[0] = m; ms[1] = ssum;
                    red2[0] = m;
                }
            }
            __syncthreads();
            if (arow < B) {
                float c = 0.f;
                for (int j = 0; j < Tq; ++j) c = fmaf(p_s[j], vals_s[j * ENC + tid], c);
                ws[P.ws.att_ctx + ((int64_t)arow * 4 + aq) * ENC + tid] = c;
            }
            // p_s stays valid until the next K_ATT stage of this CTA (finalised in K_AL)
            continue;
        }

        const StageDesc& d = P.st[s];
        const int K0 = d.K0, K1 = d.K1, K = K0 + K1, NC = d.NC;

        // =============== stage inputs: rows rg*8.. of the two sources -> act_s ===============
        {
            const float* src0 = nullptr; int ld0 = 0;
            const float* src1 = nullptr; int ld1 = 0;
            switch (s) {
                case K_P1: break;   // special loader below
                case K_P2: src0 = ws + P.ws.p1; ld0 = 256; break;
                case K_IN: src0 = ws + P.ws.p2; ld0 = 128; src1 = ws + P.ws.attn; ld1 = AU; break;
                case K_G1: src0 = ws + P.ws.z; ld0 = U; src1 = ws + P.ws.h[0]; ld1 = U; break;
                case K_C1: src0 = ws + P.ws.z; ld0 = U; src1 = ws + P.ws.rh; ld1 = U; break;
                case K_G2: src0 = ws + P.ws.h[0]; ld0 = U; src1 = ws + P.ws.h[1]; ld1 = U; break;
                case K_C2: src0 = ws + P.ws.h[0]; ld0 = U; src1 = ws + P.ws.rh; ld1 = U; break;
                case K_G3: src0 = ws + P.ws.h[1]; ld0 = U; src1 = ws + P.ws.h[2]; ld1 = U; break;
                case K_C3: src0 = ws + P.ws.h[1]; ld0 = U; src1 = ws + P.ws.rh; ld1 = U; break;
                case K_OUT: src0 = ws + P.ws.s; ld0 = U; break;
                case K_Q:  src0 = ws + P.ws.ybuf; ld0 = 512; break;
                case K_AL: src0 = ws + P.ws.ybuf; ld0 = 512; break;   // ctx handled below
            }
            if (s == K_P1) {
                // decoder input for step t, last mel frame of the r-group (tacotron.py:66-67; helpers A.8-A.10)
                for (int i = tid; i < RPG * (MF / 4); i += NTHR) {
                    int r = i / (MF / 4), c4 = (i % (MF / 4)) * 4;
                    int row = rg * RPG + r;
                    float4 v = make_float4(0, 0, 0, 0);
                    if (row < B) {
                        bool from_y;
                        if (A.mode == TACO_DEC_INFER) from_y = true;
                        else if (A.mode == TACO_DEC_TEACHER) from_y = false;
                        else from_y = (t > 0) && (A.sample_mask[(int64_t)(t - 1) * B + row] != 0);
                        if (from_y) {
                            if (t > 0) v = ldcg4(ws + P.ws.ybuf + (int64_t)row * 512 + (OUT - MF) + c4);
                        } else {
                            v = __ldg(reinterpret_cast<const float4*>(A.mel + ((int64_t)row * T + t) * OUT + (OUT - MF) + c4));
                        }
                    }
                    *reinterpret_cast<float4*>(act_s + r * ACT_LD + c4) = v;
                }
            } else {
                const int k04 = K0 / 4;
                for (int i = tid; i < RPG * k04; i += NTHR) {
                    int r = i / k04, c4 = (i % k04) * 4;
                    float4 v = ldcg4(src0 + (int64_t)(rg * RPG + r) * ld0 + c4);
                    *reinterpret_cast<float4*>(act_s + r * ACT_LD + c4) = v;
                }
                if (s == K_AL) {
                    // ctx = flash-style merge of the four quarter partials (values-weighted sums)
                    for (int i = tid; i < RPG * (ENC / 4); i += NTHR) {
                        int r = i / (ENC / 4), c4 = (i % (ENC / 4)) * 4;
                        int row = rg * RPG + r;
                        float4 acc = make_float4(0, 0, 0, 0);
                        if (row < B) {
                            const float* ms = ws + P.ws.att_ms + (int64_t)row * 8;
                            float m[4], sq[4], M = -INFINITY;
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) { m[qd] = __ldcg(ms + 2 * qd); sq[qd] = __ldcg(ms + 2 * qd + 1); M = fmaxf(M, m[qd]); }
                            float S = 0.f, w[4];
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) { w[qd] = (m[qd] == -INFINITY) ? 0.f : __expf(m[qd] - M); S += w[qd] * sq[qd]; }
                            const float inv = 1.0f / S;
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) {
                                float4 c = ldcg4(ws + P.ws.att_ctx + ((int64_t)row * 4 + qd) * ENC + c4);
                                float ww = w[qd] * inv;
                                acc.x = fmaf(ww, c.x, acc.x); acc.y = fmaf(ww, c.y, acc.y);
                                acc.z = fmaf(ww, c.z, acc.z); acc.w = fmaf(ww, c.w, acc.w);
                            }
                        }
                        *reinterpret_cast<float4*>(act_s + r * ACT_LD + K0 + c4) = acc;
                    }
                    // finalise this CTA's slice of the alignments: a_j = p_j * exp(m_q - M) / S
                    if (arow < B) {
                        const float* ms = ws + P.ws.att_ms + (int64_t)arow * 8;
                        float m[4], sq[4], M = -INFINITY;
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) { m[qd] = __ldcg(ms + 2 * qd); sq[qd] = __ldcg(ms + 2 * qd + 1); M = fmaxf(M, m[qd]); }
                        float S = 0.f;
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) S += ((m[qd] == -INFINITY) ? 0.f : __expf(m[qd] - M)) * sq[qd];
                        const float sc = ((m[aq] == -INFINITY) ? 0.f : __expf(m[aq] - M)) / S;
                        const float* p_s = small_s + 512 + 64;
                        for (int j = tid; j < Tq; j += NTHR)
                            A.align[((int64_t)arow * T + t) * Tx + aq * Tq + j] = p_s[j] * sc;
                    }
                } else if (K1 > 0) {
                    const int k14 = K1 / 4;
                    for (int i = tid; i < RPG * k14; i += NTHR) {
                        int r = i / k14, c4 = (i % k14) * 4;
                        float4 v = ldcg4(src1 + (int64_t)(rg * RPG + r) * ld1 + c4);
                        *reinterpret_cast<float4*>(act_s + r * ACT_LD + K0 + c4) = v;
                    }
                }
            }
        }
        // ---- this stage's weight slice ----
        const float* Wsl;
        if (d.res_off >= 0) {
            Wsl = res_s + d.res_off;
        } else {
            int buf = n_consumed & 1;
            mbar_wait(&wbar[buf], (n_consumed >> 1) & 1);
            Wsl = stream_s + buf * STREAM_FLOATS;
            ++n_consumed;
        }
        __syncthreads();

        // =============== partial dot products: thread = (row r, column quad cq, k-slice) ===============
        const int r = lane & 7, g = lane >> 3;
        const int ksw = 16 / NC;                 // k-slices inside a warp: NC=16 ->1, 8 ->2, 4 ->4
        const int cq = g % (NC / 4);
        const int ks = g / (NC / 4);
        const int nsl = 8 * ksw;
        const int slice = warp * ksw + ks;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* xr = act_s + r * ACT_LD;
        for (int c = slice; c < K / 4; c += nsl) {
            const float4 x = *reinterpret_cast<const float4*>(xr + 4 * c);
            const float* wp = Wsl + (4 * c) * NC + cq * 4;
            const float4 w0 = *reinterpret_cast<const float4*>(wp);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + NC);
            const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * NC);
            const float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * NC);
            acc[0] = fmaf(x.x, w0.x, acc[0]); acc[1] = fmaf(x.x, w0.y, acc[1]); acc[2] = fmaf(x.x, w0.z, acc[2]); acc[3] = fmaf(x.x, w0.w, acc[3]);
            acc[0] = fmaf(x.y, w1.x, acc[0]); acc[1] = fmaf(x.y, w1.y, acc[1]); acc[2] = fmaf(x.y, w1.z, acc[2]); acc[3] = fmaf(x.y, w1.w, acc[3]);
            acc[0] = fmaf(x.z, w2.x, acc[0]); acc[1] = fmaf(x.z, w2.y, acc[1]); acc[2] = fmaf(x.z, w2.z, acc[2]); acc[3] = fmaf(x.z, w2.w, acc[3]);
            acc[0] = fmaf(x.w, w3.x, acc[0]); acc[1] = fmaf(x.w, w3.y, acc[1]); acc[2] = fmaf(x.w, w3.z, acc[2]); acc[3] = fmaf(x.w, w3.w, acc[3]);
        }
        // reduce the in-warp k-slices (lanes differing in the ks bits of g)
        if (NC == 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
        } else if (NC == 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8); acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16); }
        }
        *reinterpret_cast<float4*>(red_s + (warp * 32 + lane) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        __syncthreads();

        // =============== cross-warp reduction + stage epilogue: one thread per output ===============
        const int nout = RPG * NC;            // valid lanes = 8*NC/4, x4 columns
        if (tid < nout) {
            const int lo = tid >> 2, j = tid & 3;       // lane index holding (r, cq), column j of the quad
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red_s[(w * 32 + lo) * 4 + j];
            const int rr = lo & 7, cqq = lo >> 3;
            const int row = rg * RPG + rr;
            const int col = cs * NC + cqq * 4 + j;
            const int64_t ro = (int64_t)row;
            switch (s) {
                case K_P1: {
                    v = fmaxf(v + __ldg(P.pre_b1 + col), 0.f);
                    if (A.keep1 && row < B) v = A.keep1[((int64_t)t * B + row) * 256 + col] ? v * A.keep_scale : 0.f;
                    ws[P.ws.p1 + ro * 256 + col] = v;
                } break;
                case K_P2: {
                    v = fmaxf(v + __ldg(P.pre_b2 + col), 0.f);
                    if (A.keep2 && row < B) v = A.keep2[((int64_t)t * B + row) * 128 + col] ? v * A.keep_scale : 0.f;
                    ws[P.ws.p2 + ro * 128 + col] = v;
                } break;
                case K_IN: ws[P.ws.z + ro * U + col] = v + __ldg(P.in_b + col); break;
                case K_G1: case K_G2: case K_G3: {
                    const int gi = (s - K_G1) / 2;
                    float gte = sigmoidf_acc(v + __ldg(P.gru_bg[gi] + col));
                    if (col < U) ws[P.ws.rh + ro * U + col] = gte * __ldcg(ws + P.ws.h[gi] + ro * U + col);
                    else ws[P.ws.u + ro * U + (col - U)] = gte;
                } break;
                case K_C1: case K_C2: case K_C3: {
                    const int gi = (s - K_C1) / 2;
                    float c = tanhf_acc(v + __ldg(P.gru_bc[gi] + col));
                    float uu = __ldcg(ws + P.ws.u + ro * U + col);
                    float hold = __ldcg(ws + P.ws.h[gi] + ro * U + col);
                    float hn = uu * hold + (1.0f - uu) * c;
                    ws[P.ws.h[gi] + ro * U + col] = hn;
                    if (gi == 2) ws[P.ws.s + ro * U + col] = __ldcg(ws + P.ws.z + ro * U + col) + hn;
                } break;
                case K_OUT: {
                    if (col < OUT) {
                        float y = v + __ldg(P.out_b + col);
                        ws[P.ws.ybuf + ro * 512 + col] = y;
                        if (row < B) A.y[((int64_t)row * T + t) * OUT + col] = y;
                    }
                } break;
                case K_Q: ws[P.ws.q + ro * AU + col] = v; break;
                case K_AL: ws[P.ws.attn + ro * AU + col] = v; break;
            }
        }
        // the trailing __syncthreads of grid_barrier (next iteration) orders these smem reads before reuse
    }
}

// ---- packing: TF [K][N] -> per-slice [cs][K][NC] ----
__global__ void pack_stage_kernel(const float* __restrict__ W, int K, int N, int NC, float* __restrict__ dst) {
    int64_t total = (int64_t)NS * K * NC;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(i % NC);
        int64_t rest = i / NC;
        int k = (int)(rest % K);
        int cs = (int)(rest / K);
        int col = cs * NC + j;
        dst[i] = (col < N) ? W[(int64_t)k * N + col] : 0.0f;
    }
}

void build_stage_table(int r, StageDesc* st, int64_t* total_floats) {
    const int OUT = MF * r;
    int64_t off = 0;
    for (int s = 0; s < NSTAGE; ++s) {
        int K0, K1, N, NC;
        stage_dims(s, OUT, K0, K1, N, NC);
        st[s].K0 = K0; st[s].K1 = K1; st[s].N = N; st[s].NC = NC; st[s].w_off = off; st[s].res_off = -1;
        off += (int64_t)NS * (K0 + K1) * NC;
    }
    *total_floats = off;
}

void build_ws_layout(DecLayout* L) {
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += (n + 31) / 32 * 32; return r; };
    L->p1 = take(BPAD * 256); L->p2 = take(BPAD * 128); L->attn = take(BPAD * AU); L->z = take(BPAD * U);
    for (int i = 0; i < 3; ++i) L->h[i] = take(BPAD * U);
    L->rh = take(BPAD * U); L->u = take(BPAD * U); L->s = take(BPAD * U); L->ybuf = take(BPAD * 512);
    L->q = take(BPAD * AU); L->att_ms = take(BPAD * 8); L->att_ctx = take(BPAD * 4 * ENC); L->ctr = take(64);
    L->total = o;
}

}  // namespace

extern "C" size_t taco_decoder_packed_bytes(int r) {
    StageDesc st[NSTAGE]; int64_t tot;
    build_stage_table(r, st, &tot);
    return (size_t)tot * 4;
}

extern "C" size_t taco_decoder_workspace_bytes(int B, int Tx, int T, int r) {
    (void)B; (void)Tx; (void)T; (void)r;
    DecLayout L; build_ws_layout(&L);
    return (size_t)L.total * 4;
}

extern "C" int taco_decoder_pack(const taco_decoder_weights* w, int r, float* packed, void* stream) {
    TACO_CHECK(w && packed, "taco_decoder_pack: NULL");
    TACO_CHECK(r >= 1 && MF * r <= 512, "taco_decoder_pack: r=%d out of range (80r <= 512)", r);
    StageDesc st[NSTAGE]; int64_t tot;
    build_stage_table(r, st, &tot);
    const float* src[NSTAGE] = {w->pre_W1, w->pre_W2, w->in_W, w->gru_Wg[0], w->gru_Wc[0], w->gru_Wg[1], w->gru_Wc[1],
                                w->gru_Wg[2], w->gru_Wc[2], w->out_W, w->att_Wq, nullptr, w->att_Wa};
    for (int s = 0; s < NSTAGE; ++s) {
        if (s == K_ATT) continue;
        TACO_CHECK(src[s] != nullptr, "taco_decoder_pack: weight %d is NULL", s);
        int K = st[s].K0 + st[s].K1;
        int64_t total = (int64_t)NS * K * st[s].NC;
        int blocks = (int)((total + 255) / 256);
        pack_stage_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src[s], K, st[s].N, st[s].NC, packed + st[s].w_off);
        TACO_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int taco_decoder_fwd(const taco_decoder_args* a, void* stream) {
    TACO_CHECK(a, "taco_decoder_fwd: NULL args");
    TACO_CHECK(a->weights != nullptr, "taco_decoder_fwd: weights (biases, attention_v) is NULL");
    const taco_decoder_weights& g_dec_w = *a->weights;
    TACO_CHECK(a->B >= 1 && a->B <= BPAD, "taco_decoder_fwd: B=%d must be in [1,%d] per launch", a->B, BPAD);
    TACO_CHECK(a->T >= 1, "taco_decoder_fwd: T=%d", a->T);
    TACO_CHECK(a->Tx >= 4 && (a->Tx % 4) == 0 && a->Tx <= 256, "taco_decoder_fwd: Tx=%d must be a multiple of 4, <= 256", a->Tx);
    TACO_CHECK(a->r >= 1 && MF * a->r + ENC <= ACT_LD - 4, "taco_decoder_fwd: r=%d unsupported (80r + 256 must be <= %d)", a->r, ACT_LD - 4);
    TACO_CHECK(a->packed && a->keys && a->values && a->text_length && a->y && a->align && a->workspace, "taco_decoder_fwd: NULL pointer");
    if (a->mode != TACO_DEC_INFER) TACO_CHECK(a->mel != nullptr, "taco_decoder_fwd: teacher/sched mode needs mel");
    if (a->mode == TACO_DEC_SCHED) TACO_CHECK(a->sample_mask != nullptr, "taco_decoder_fwd: sched mode needs sample_mask");
    TACO_CHECK((a->keep1 == nullptr) == (a->keep2 == nullptr), "taco_decoder_fwd: keep1/keep2 must both be set or both NULL");
    cudaStream_t st = (cudaStream_t)stream;

    DecParams P;
    memset(&P, 0, sizeof(P));
    int64_t tot;
    build_stage_table(a->r, P.st, &tot);
    build_ws_layout(&P.ws);
    P.a = *a;
    P.OUT = MF * a->r;
    P.Tq = a->Tx / 4;
    TACO_CHECK(P.Tq <= 64, "taco_decoder_fwd: Tx/4 = %d > 64", P.Tq);
    P.pre_b1 = g_dec_w.pre_b1; P.pre_b2 = g_dec_w.pre_b2; P.in_b = g_dec_w.in_b; P.out_b = g_dec_w.out_b; P.att_v = g_dec_w.att_v;
    for (int i = 0; i < 3; ++i) { P.gru_bg[i] = g_dec_w.gru_bg[i]; P.gru_bc[i] = g_dec_w.gru_bc[i]; }

    // shared-memory plan
    int off = 64 + RPG * ACT_LD + 8 * 32 * 4 + 1024;
    off = (off + 31) / 32 * 32;
    P.smem_kv_off = off;
    off += P.Tq * KV_LD + P.Tq * ENC;
    off = (off + 31) / 32 * 32;
    P.smem_stream_off = off;
    off += 2 * STREAM_FLOATS;
    P.smem_res_off = off;
    const int max_floats = (227 * 1024) / 4;
    int budget = max_floats - off;
    // greedy residency: biggest per-step traffic first (GRU gates, candidates, attention layer, ...)
    const int order[] = {K_G1, K_G2, K_G3, K_C1, K_C2, K_C3, K_AL, K_IN, K_OUT, K_Q, K_P1, K_P2};
    int res = 0;
    for (int i = 0; i < 12; ++i) {
        StageDesc& d = P.st[order[i]];
        int fl = (d.K0 + d.K1) * d.NC;
        if (fl <= budget - res) { d.res_off = res; res += fl; }
    }
    P.res_floats = res;
    P.smem_total_floats = off + res;
    const size_t smem_bytes = (size_t)P.smem_total_floats * 4;

    static bool configured = false;
    if (!configured) {
        TACO_CUDA(cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    TACO_CUDA(cudaMemsetAsync(a->workspace, 0, (size_t)P.ws.total * 4, st));
    void* args[] = {(void*)&P};
    TACO_CUDA(cudaLaunchCooperativeKernel((void*)decoder_kernel, dim3(NCTA), dim3(NTHR), args, smem_bytes, st));
    ++g_taco_launches;
    return 0;
}
