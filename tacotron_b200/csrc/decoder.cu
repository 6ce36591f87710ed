// decoder.cu -- the whole autoregressive attention decoder as ONE persistent dataflow kernel.
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
// Design v2 (B <= 32 utterances per launch), driven by the v1 measurements (58 us / step, all of it
// grid-barrier latency + shared-memory operand traffic):
//   * grid = 128 CTAs x 256 threads, co-resident (cooperative launch), alive for all T steps.
//     Dense stage [32 x K].[K x N]: CTA (rg, cs) owns rows 8rg..8rg+7 and the cs-th of 32 column slices.
//   * NO grid barrier.  Every exchanged activation is a 64-bit word {fp32 value, step tag} written with
//     one st.b64 and read with polling 128-bit volatile loads: data and flag travel together ("LL"
//     protocol), so a stage boundary costs one L2 write->read latency, no fences, no atomics.  A buffer
//     is re-written only one full step later, which the dependency chain itself guarantees to be after
//     all of its readers have consumed it.
//   * compute mapping: thread tile 8 rows x 4 columns, the 32 lanes of a warp split K in 16-byte
//     chunks (12 LDS.128 per 128 FMA instead of 5 per 16 in v1 -- shared-memory->register bandwidth
//     was the limiter), lane partials combined by a 32->1 halving butterfly (31 shuffles), warps of
//     the same column quad by a 1 KB shared-memory exchange.
//   * GRU gate columns are permuted so that CTA cs owns r and u of the SAME 8 hidden units: u, the
//     previous state and z never leave the CTA.
//   * K x NC weight slices are contiguous in the packed buffer (taco_decoder_pack) and are either
//     RESIDENT in shared memory or streamed from L2 by cp.async.bulk + mbarrier, double buffered,
//     issued one stage ahead.
//   * attention: CTA (utterance, quarter of Tx) keeps its keys/values slice in shared memory for all
//     steps; scores + partial softmax + partial context per quarter, flash-style merge by the consumer.
//   * %globaltimer stamp per step for the decoder-step latency metric.
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
constexpr int ACT_LD = 660;    // 656 + 4 floats: row stride of the staged activations
constexpr int KV_LD = 260;     // padded row stride for keys in smem
constexpr int STREAM_FLOATS = 512 * 16;   // largest weight slice (K=512, NC=16)
constexpr int YLD = 512;       // row stride (elements) of the y exchange buffer

enum StageKind { K_P1 = 0, K_P2, K_IN, K_G1, K_C1, K_G2, K_C2, K_G3, K_C3, K_OUT, K_Q, K_ATT, K_AL };

struct StageDesc {
    int K0, K1;        // K = K0 + K1 (two concatenated sources)
    int N;             // true output width
    int NC;            // columns per slice (4, 8 or 16); padded width = NC*32
    int64_t w_off;     // float offset of slice 0 in the packed buffer
    int res_off;       // float offset inside the resident smem region, or -1 = streamed
};

struct DecLayout {     // workspace offsets in 64-bit LL words; every buffer is [32][ld]
    int64_t p1, p2, attn, z, h[3], rh[3], s, ybuf, q, att_ms, att_ctx;
    int64_t total;
};

struct DecParams {
    StageDesc st[NSTAGE];
    DecLayout ws;
    taco_decoder_args a;
    const float *pre_b1, *pre_b2, *in_b, *gru_bg[3], *gru_bc[3], *out_b, *att_v;
    int OUT;           // 80*r
    int Tq;            // Tx/4
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

// global column computed by local column j of slice cs.  GRU gates: slice cs owns r (j < 8) and u
// (j >= 8) of hidden units 8cs..8cs+7.
__host__ __device__ inline int stage_col(int kind, int cs, int j, int NC) {
    if (kind == K_G1 || kind == K_G2 || kind == K_G3) return (j < 8) ? (8 * cs + j) : (U + 8 * cs + (j - 8));
    return cs * NC + j;
}

// ---------------------------------------------------------------------------------------------
// LL words: {value, tag}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint64_t* p, float v, uint32_t tag) {
    const uint64_t w = (uint64_t)__float_as_uint(v) | ((uint64_t)tag << 32);
    asm volatile("st.global.cg.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load2(const uint64_t* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool ll_ok(const ulonglong2& v, uint32_t tag) {
    return (uint32_t)(v.x >> 32) == tag && (uint32_t)(v.y >> 32) == tag;
}
// Wait for two consecutive LL words carrying `tag`; bounded so that a protocol bug traps instead of hanging.
__device__ __forceinline__ float2 ll_wait2(const uint64_t* p, uint32_t tag) {
    ulonglong2 v = ll_load2(p);
    uint32_t spins = 0;
    while (!ll_ok(v, tag)) {
        if (++spins > (1u << 24)) __trap();
        v = ll_load2(p);
    }
    return make_float2(__uint_as_float((uint32_t)v.x), __uint_as_float((uint32_t)v.y));
}

// rows [row0, row0+8) x W columns of an LL buffer (row stride ld words) -> act_s[r][dst0 + c]
__device__ __forceinline__ void ingest(float* act_s, int dst0, const uint64_t* buf, int ld, int row0, int W, uint32_t tag) {
    const int wp = W >> 1;                 // pairs per row
    const int npairs = RPG * wp;
    for (int base = threadIdx.x; base < npairs; base += NTHR * 4) {
        ulonglong2 v[4];
        const uint64_t* ptr[4];
        int r[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = base + u * NTHR;
            r[u] = p / wp; c[u] = (p - r[u] * wp) * 2;
            ptr[u] = buf + (int64_t)(row0 + r[u]) * ld + c[u];
            if (p < npairs) v[u] = ll_load2(ptr[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = base + u * NTHR;
            if (p < npairs) {
                uint32_t spins = 0;
                while (!ll_ok(v[u], tag)) {
                    if (++spins > (1u << 24)) __trap();
                    v[u] = ll_load2(ptr[u]);
                }
                *reinterpret_cast<float2*>(act_s + r[u] * ACT_LD + dst0 + c[u]) =
                    make_float2(__uint_as_float((uint32_t)v[u].x), __uint_as_float((uint32_t)v[u].y));
            }
        }
    }
}
__device__ __forceinline__ void ingest_zero(float* act_s, int dst0, int W) {
    for (int i = threadIdx.x; i < RPG * W; i += NTHR) act_s[(i / W) * ACT_LD + dst0 + (i % W)] = 0.f;
}

// 32 -> 1 halving butterfly: on return lane l holds the warp-wide sum of element l.
__device__ __forceinline__ float butterfly32(float (&p)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < n) {
                const float send = hi ? p[j] : p[j + n];
                const float keep = hi ? p[j + n] : p[j];
                p[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return p[0];
}

__global__ void __launch_bounds__(NTHR, 1) decoder_kernel(const DecParams P) {
    extern __shared__ __align__(16) float smem[];
    // smem map (floats): [0,64) mbarriers | act_s 8*ACT_LD | part_s 8*32 | loc 512 | small 1024 | kv | stream 2x | resident
    uint64_t* wbar = reinterpret_cast<uint64_t*>(smem);          // [0],[1] stream buffers, [2] resident preload
    float* act_s = smem + 64;
    float* part_s = act_s + RPG * ACT_LD;                         // [8 warps][32]
    float* loc_s = part_s + 8 * 32;                               // h_loc[3][64] | u_loc[64] | z_loc[64]
    float* small_s = loc_s + 512;                                 // q_s[256] | v_s[256] | e_s[64] | p_s[64] | misc
    float* keys_s = smem + P.smem_kv_off;
    float* vals_s = keys_s + P.Tq * KV_LD;
    float* res_s = smem + P.smem_res_off;
    float* stream_s = smem + P.smem_stream_off;
    float* h_loc = loc_s;             // [3][8][8]
    float* u_loc = loc_s + 192;       // [8][8]
    float* z_loc = loc_s + 256;       // [8][8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x;
    const int rg = cta & 3, cs = cta >> 2;             // dense stages
    const int arow = cta >> 2, aq = cta & 3;           // attention stage: utterance, quarter
    const taco_decoder_args& A = P.a;
    uint64_t* ws = reinterpret_cast<uint64_t*>(A.workspace);
    const int B = A.B, T = A.T, OUT = P.OUT, Tq = P.Tq, Tx = A.Tx;
    const int row0 = rg * RPG;

    if (tid == 0) {
        mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); mbar_init(&wbar[2], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 512; i += NTHR) loc_s[i] = 0.f;          // zero initial GRU states (cell.zero_state)
    __syncthreads();

    // ---- one-time preload: resident weight slices (bulk copies) ----
    if (tid == 0) {
        uint32_t bytes = 0;
        for (int s = 0; s < NSTAGE; ++s) {
            const StageDesc& d = P.st[s];
            if (s == K_ATT || d.res_off < 0) continue;
            bytes += (uint32_t)((d.K0 + d.K1) * d.NC * 4);
        }
        if (bytes) {
            mbar_arrive_expect_tx(&wbar[2], bytes);
            for (int s = 0; s < NSTAGE; ++s) {
                const StageDesc& d = P.st[s];
                if (s == K_ATT || d.res_off < 0) continue;
                const uint32_t nb = (uint32_t)((d.K0 + d.K1) * d.NC * 4);
                bulk_load(res_s + d.res_off, A.packed + d.w_off + (int64_t)cs * (d.K0 + d.K1) * d.NC, nb, &wbar[2]);
            }
        } else {
            mbar_arrive(&wbar[2]);
        }
    }
    // keys / values slice of (arow, aq) -> smem (padded rows), zero for utterances >= B
    for (int i = tid; i < Tq * (ENC / 4); i += NTHR) {
        const int j = i / (ENC / 4), d4 = (i % (ENC / 4)) * 4;
        float4 kk = make_float4(0, 0, 0, 0), vv = kk;
        if (arow < B) {
            const int64_t g = ((int64_t)arow * Tx + aq * Tq + j) * ENC + d4;
            kk = __ldg(reinterpret_cast<const float4*>(A.keys + g));
            vv = __ldg(reinterpret_cast<const float4*>(A.values + g));
        }
        *reinterpret_cast<float4*>(keys_s + j * KV_LD + d4) = kk;
        *reinterpret_cast<float4*>(vals_s + j * ENC + d4) = vv;
    }
    for (int i = tid; i < AU; i += NTHR) small_s[256 + i] = __ldg(P.att_v + i);
    const int my_len = (arow < B) ? A.text_length[arow] : 0;
    mbar_wait(&wbar[2], 0);
    __syncthreads();

    // streamed-slice bookkeeping: issue order == consume order; buffer = (index) & 1.  At most two slices
    // are in flight.  A buffer is re-filled only at the top of an iteration, when every thread has passed
    // the post-compute __syncthreads of the iteration that last read it.
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

    for (int it = 0; it < total_iters; ++it) {
        const int t = it / NSTAGE;
        const int s = it - t * NSTAGE;
        const uint32_t tag_now = (uint32_t)t + 1;      // values produced during step t
        const uint32_t tag_prev = (uint32_t)t;         // values produced during step t-1 (t = 0: initial zeros)
        pump();
        if (s == 0 && cta == 0 && tid == 0 && A.step_ns) A.step_ns[t] = globaltimer_ns();

        if (s == K_ATT) {
            // =============== attention scores / partial softmax / partial context ===============
            float* q_s = small_s;             // [256]
            float* v_s = small_s + 256;       // [256] (loaded once)
            float* e_s = small_s + 512;       // [Tq]
            float* p_s = small_s + 512 + 64;  // [Tq]  (kept until the K_AL stage)
            if (arow < B) {
                if (tid < AU / 2) {
                    const float2 qq = ll_wait2(ws + P.ws.q + (int64_t)arow * AU + 2 * tid, tag_now);
                    q_s[2 * tid] = qq.x; q_s[2 * tid + 1] = qq.y;
                }
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
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int d0 = i * 32 + dsl * 4;
                            const float4 kk = *reinterpret_cast<const float4*>(keys_s + j * KV_LD + d0);
                            const float4 qq = *reinterpret_cast<const float4*>(q_s + d0);
                            const float4 vv = *reinterpret_cast<const float4*>(v_s + d0);
                            acc = fmaf(vv.x, tanhf_acc(kk.x + qq.x), acc);
                            acc = fmaf(vv.y, tanhf_acc(kk.y + qq.y), acc);
                            acc = fmaf(vv.z, tanhf_acc(kk.z + qq.z), acc);
                            acc = fmaf(vv.w, tanhf_acc(kk.w + qq.w), acc);
                        }
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
                    const float p = (m == -INFINITY) ? 0.f : __expf(e_s[j] - m);
                    p_s[j] = p;
                    ssum += p;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                if (lane == 0) {
                    uint64_t* ms = ws + P.ws.att_ms + ((int64_t)arow * 4 + aq) * 2;
                    ll_store(ms, m, tag_now);
                    ll_store(ms + 1, ssum, tag_now);
                }
            }
            __syncthreads();
            if (arow < B) {
                float c = 0.f;
                for (int j = 0; j < Tq; ++j) c = fmaf(p_s[j], vals_s[j * ENC + tid], c);
                ll_store(ws + P.ws.att_ctx + ((int64_t)arow * 4 + aq) * ENC + tid, c, tag_now);
            }
            continue;
        }

        const StageDesc& d = P.st[s];
        const int K0 = d.K0, K1 = d.K1, K = K0 + K1, NC = d.NC;

        // =============== stage inputs: rows row0..row0+7 of the sources -> act_s (polling LL loads) ===============
        switch (s) {
            case K_P1: {
                // decoder input for step t, last mel frame of the r-group (tacotron.py:66-67; helpers A.8-A.10)
                for (int i = tid; i < RPG * (MF / 2); i += NTHR) {
                    const int r = i / (MF / 2), c2 = (i % (MF / 2)) * 2;
                    const int row = row0 + r;
                    float2 v = make_float2(0.f, 0.f);
                    if (row < B) {
                        bool from_y;
                        if (A.mode == TACO_DEC_INFER) from_y = true;
                        else if (A.mode == TACO_DEC_TEACHER) from_y = false;
                        else from_y = (t > 0) && (A.sample_mask[(int64_t)(t - 1) * B + row] != 0);
                        if (from_y) {
                            if (t > 0) v = ll_wait2(ws + P.ws.ybuf + (int64_t)row * YLD + (OUT - MF) + c2, tag_prev);
                        } else {
                            v = __ldg(reinterpret_cast<const float2*>(A.mel + ((int64_t)row * T + t) * OUT + (OUT - MF) + c2));
                        }
                    }
                    *reinterpret_cast<float2*>(act_s + r * ACT_LD + c2) = v;
                }
            } break;
            case K_P2: ingest(act_s, 0, ws + P.ws.p1, 256, row0, 256, tag_now); break;
            case K_IN:
                ingest(act_s, 0, ws + P.ws.p2, 128, row0, 128, tag_now);
                if (t > 0) ingest(act_s, 128, ws + P.ws.attn, AU, row0, AU, tag_prev); else ingest_zero(act_s, 128, AU);
                break;
            case K_G1: case K_G2: case K_G3: {
                const int gi = (s - K_G1) / 2;
                const uint64_t* xsrc = (gi == 0) ? ws + P.ws.z : ws + P.ws.h[gi - 1];
                ingest(act_s, 0, xsrc, U, row0, U, tag_now);
                if (t > 0) ingest(act_s, U, ws + P.ws.h[gi], U, row0, U, tag_prev); else ingest_zero(act_s, U, U);
            } break;
            case K_C1: case K_C2: case K_C3: {
                const int gi = (s - K_C1) / 2;          // x part is still in act_s[.., 0:256) from the gate stage
                ingest(act_s, U, ws + P.ws.rh[gi], U, row0, U, tag_now);
            } break;
            case K_OUT: ingest(act_s, 0, ws + P.ws.s, U, row0, U, tag_now); break;
            case K_Q: ingest(act_s, 0, ws + P.ws.ybuf, YLD, row0, OUT, tag_now); break;
            case K_AL: {
                // y is still in act_s[.., 0:OUT) from the K_Q stage (K_ATT does not touch act_s).
                // ctx = flash-style merge of the four quarter partials.  First the (max, sum) pairs of the 8 rows
                // of this row group and of this CTA's own attention row -> shared memory (36 threads poll).
                float* ms_s = small_s + 704;          // [9][4][2]: rows 0..7 = row group, row 8 = arow
                if (tid < 36) {
                    const int rsel = tid >> 2, qd = tid & 3;
                    const int row = (rsel < 8) ? row0 + rsel : arow;
                    float2 x = make_float2(-INFINITY, 0.f);
                    if (row < B) x = ll_wait2(ws + P.ws.att_ms + (int64_t)row * 8 + 2 * qd, tag_now);
                    ms_s[tid * 2] = x.x; ms_s[tid * 2 + 1] = x.y;
                }
                __syncthreads();
                for (int i = tid; i < RPG * (ENC / 2); i += NTHR) {
                    const int r = i / (ENC / 2), c2 = (i % (ENC / 2)) * 2;
                    const int row = row0 + r;
                    float2 acc = make_float2(0.f, 0.f);
                    if (row < B) {
                        float M = -INFINITY;
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) M = fmaxf(M, ms_s[(r * 4 + qd) * 2]);
                        float S = 0.f, w[4];
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const float m = ms_s[(r * 4 + qd) * 2];
                            w[qd] = (m == -INFINITY) ? 0.f : __expf(m - M);
                            S += w[qd] * ms_s[(r * 4 + qd) * 2 + 1];
                        }
                        const float inv = 1.0f / S;
                        float2 c[4];
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            c[qd] = ll_wait2(ws + P.ws.att_ctx + ((int64_t)row * 4 + qd) * ENC + c2, tag_now);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const float ww = w[qd] * inv;
                            acc.x = fmaf(ww, c[qd].x, acc.x); acc.y = fmaf(ww, c[qd].y, acc.y);
                        }
                    }
                    *reinterpret_cast<float2*>(act_s + r * ACT_LD + K0 + c2) = acc;
                }
                // finalise this CTA's slice of the alignments: a_j = p_j * exp(m_q - M) / S
                if (arow < B) {
                    float M = -INFINITY;
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) M = fmaxf(M, ms_s[(32 + qd) * 2]);
                    float S = 0.f;
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float m = ms_s[(32 + qd) * 2];
                        S += ((m == -INFINITY) ? 0.f : __expf(m - M)) * ms_s[(32 + qd) * 2 + 1];
                    }
                    const float mq = ms_s[(32 + aq) * 2];
                    const float sc = ((mq == -INFINITY) ? 0.f : __expf(mq - M)) / S;
                    const float* p_s = small_s + 512 + 64;
                    for (int j = tid; j < Tq; j += NTHR)
                        A.align[((int64_t)arow * T + t) * Tx + aq * Tq + j] = p_s[j] * sc;
                }
            } break;
        }
        // ---- this stage's weight slice ----
        const float* Wsl;
        if (d.res_off >= 0) {
            Wsl = res_s + d.res_off;
        } else {
            const int buf = n_consumed & 1;
            mbar_wait(&wbar[buf], (n_consumed >> 1) & 1);
            Wsl = stream_s + buf * STREAM_FLOATS;
            ++n_consumed;
        }
        __syncthreads();

        // =============== partial products: thread tile = 8 rows x 4 columns, lanes split K ===============
        const int NQ = NC >> 2;                 // column quads in the slice: 1, 2 or 4
        const int nkp = 8 / NQ;                 // warps (k parts) per quad
        const int cq = warp % NQ;
        const int kp = warp / NQ;
        const int NCH = K >> 2;                 // 16-byte k chunks
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        // packed slice layout: [i = k%4][cq][chunk][4 cols]  -> lanes read consecutive 16-byte words
        const float4* W4 = reinterpret_cast<const float4*>(Wsl);
        for (int c = kp * 32 + lane; c < NCH; c += nkp * 32) {
            const float4 w0 = W4[(0 * NQ + cq) * NCH + c];
            const float4 w1 = W4[(1 * NQ + cq) * NCH + c];
            const float4 w2 = W4[(2 * NQ + cq) * NCH + c];
            const float4 w3 = W4[(3 * NQ + cq) * NCH + c];
#pragma unroll
            for (int r = 0; r < RPG; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(act_s + r * ACT_LD + 4 * c);
                acc[4 * r + 0] = fmaf(x.x, w0.x, acc[4 * r + 0]); acc[4 * r + 1] = fmaf(x.x, w0.y, acc[4 * r + 1]);
                acc[4 * r + 2] = fmaf(x.x, w0.z, acc[4 * r + 2]); acc[4 * r + 3] = fmaf(x.x, w0.w, acc[4 * r + 3]);
                acc[4 * r + 0] = fmaf(x.y, w1.x, acc[4 * r + 0]); acc[4 * r + 1] = fmaf(x.y, w1.y, acc[4 * r + 1]);
                acc[4 * r + 2] = fmaf(x.y, w1.z, acc[4 * r + 2]); acc[4 * r + 3] = fmaf(x.y, w1.w, acc[4 * r + 3]);
                acc[4 * r + 0] = fmaf(x.z, w2.x, acc[4 * r + 0]); acc[4 * r + 1] = fmaf(x.z, w2.y, acc[4 * r + 1]);
                acc[4 * r + 2] = fmaf(x.z, w2.z, acc[4 * r + 2]); acc[4 * r + 3] = fmaf(x.z, w2.w, acc[4 * r + 3]);
                acc[4 * r + 0] = fmaf(x.w, w3.x, acc[4 * r + 0]); acc[4 * r + 1] = fmaf(x.w, w3.y, acc[4 * r + 1]);
                acc[4 * r + 2] = fmaf(x.w, w3.z, acc[4 * r + 2]); acc[4 * r + 3] = fmaf(x.w, w3.w, acc[4 * r + 3]);
            }
        }
        // lane l ends up with the warp-wide sum of element l = 4*row + col
        const float wsum = butterfly32(acc, lane);
        part_s[warp * 32 + lane] = wsum;
        __syncthreads();

        // =============== cross-warp sum + stage epilogue: one thread per output ===============
        if (tid < 32 * NQ) {
            const int oq = tid >> 5, idx = tid & 31;
            float v = 0.f;
            for (int k2 = 0; k2 < nkp; ++k2) v += part_s[(k2 * NQ + oq) * 32 + idx];
            const int rr = idx >> 2;
            const int j = oq * 4 + (idx & 3);                 // local column in the slice
            const int row = row0 + rr;
            const int col = stage_col(s, cs, j, NC);
            const int64_t ro = (int64_t)row;
            switch (s) {
                case K_P1: {
                    v = fmaxf(v + __ldg(P.pre_b1 + col), 0.f);
                    if (A.keep1 && row < B) v = A.keep1[((int64_t)t * B + row) * 256 + col] ? v * A.keep_scale : 0.f;
                    ll_store(ws + P.ws.p1 + ro * 256 + col, v, tag_now);
                } break;
                case K_P2: {
                    v = fmaxf(v + __ldg(P.pre_b2 + col), 0.f);
                    if (A.keep2 && row < B) v = A.keep2[((int64_t)t * B + row) * 128 + col] ? v * A.keep_scale : 0.f;
                    ll_store(ws + P.ws.p2 + ro * 128 + col, v, tag_now);
                } break;
                case K_IN: {
                    v += __ldg(P.in_b + col);
                    z_loc[rr * 8 + j] = v;
                    ll_store(ws + P.ws.z + ro * U + col, v, tag_now);
                } break;
                case K_G1: case K_G2: case K_G3: {
                    const int gi = (s - K_G1) / 2;
                    const float g = sigmoidf_acc(v + __ldg(P.gru_bg[gi] + col));
                    if (j < 8) ll_store(ws + P.ws.rh[gi] + ro * U + col, g * h_loc[gi * 64 + rr * 8 + j], tag_now);   // r * h
                    else u_loc[rr * 8 + (j - 8)] = g;                                                                 // u stays local
                } break;
                case K_C1: case K_C2: case K_C3: {
                    const int gi = (s - K_C1) / 2;
                    const float c = tanhf_acc(v + __ldg(P.gru_bc[gi] + col));
                    const float uu = u_loc[rr * 8 + j];
                    const float hn = uu * h_loc[gi * 64 + rr * 8 + j] + (1.0f - uu) * c;
                    h_loc[gi * 64 + rr * 8 + j] = hn;
                    ll_store(ws + P.ws.h[gi] + ro * U + col, hn, tag_now);
                    if (gi == 2) ll_store(ws + P.ws.s + ro * U + col, z_loc[rr * 8 + j] + hn, tag_now);
                } break;
                case K_OUT: {
                    if (col < OUT) {
                        const float y = v + __ldg(P.out_b + col);
                        ll_store(ws + P.ws.ybuf + ro * YLD + col, y, tag_now);
                        if (row < B) A.y[((int64_t)row * T + t) * OUT + col] = y;
                    }
                } break;
                case K_Q: ll_store(ws + P.ws.q + ro * AU + col, v, tag_now); break;
                case K_AL: ll_store(ws + P.ws.attn + ro * AU + col, v, tag_now); break;
            }
        }
        // u_loc / h_loc / z_loc / part_s / act_s hazards: the next stage's post-ingest __syncthreads orders them
    }
}

// ---- packing: TF [K][N] -> per-slice [cs][i = k%4][cq][chunk = k/4][4 cols], optional gate permutation ----
__global__ void pack_stage_kernel(const float* __restrict__ W, int K, int N, int NC, int kind, float* __restrict__ dst) {
    const int NQ = NC / 4, NCH = K / 4;
    const int64_t total = (int64_t)NS * K * NC;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int jj = (int)(idx % 4);
        int64_t rest = idx / 4;
        const int c = (int)(rest % NCH); rest /= NCH;
        const int cq = (int)(rest % NQ); rest /= NQ;
        const int i = (int)(rest % 4);
        const int cs = (int)(rest / 4);
        const int k = 4 * c + i;
        const int col = stage_col(kind, cs, cq * 4 + jj, NC);
        dst[idx] = (col < N) ? W[(int64_t)k * N + col] : 0.0f;
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
    for (int i = 0; i < 3; ++i) L->rh[i] = take(BPAD * U);
    L->s = take(BPAD * U); L->ybuf = take(BPAD * YLD);
    L->q = take(BPAD * AU); L->att_ms = take(BPAD * 8); L->att_ctx = take(BPAD * 4 * ENC);
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
    return (size_t)L.total * 8;
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
        const int K = st[s].K0 + st[s].K1;
        const int64_t total = (int64_t)NS * K * st[s].NC;
        const int blocks = (int)((total + 255) / 256);
        pack_stage_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src[s], K, st[s].N, st[s].NC, s, packed + st[s].w_off);
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
    TACO_CHECK((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "taco_decoder_fwd: workspace must be 16-byte aligned");
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
    int off = 64 + RPG * ACT_LD + 8 * 32 + 512 + 1024;
    off = (off + 31) / 32 * 32;
    P.smem_kv_off = off;
    off += P.Tq * KV_LD + P.Tq * ENC;
    off = (off + 31) / 32 * 32;
    P.smem_stream_off = off;
    off += 2 * STREAM_FLOATS;
    P.smem_res_off = off;
    const int max_floats = (227 * 1024) / 4;
    const int budget = max_floats - off;
    // greedy residency: biggest per-step traffic first (GRU gates, candidates, attention layer, ...)
    const int order[] = {K_G1, K_G2, K_G3, K_C1, K_C2, K_C3, K_AL, K_IN, K_OUT, K_Q, K_P1, K_P2};
    int res = 0;
    for (int i = 0; i < 12; ++i) {
        StageDesc& d = P.st[order[i]];
        const int fl = (d.K0 + d.K1) * d.NC;
        if (fl <= budget - res) { d.res_off = res; res += fl; }
    }
    P.smem_total_floats = off + res;
    const size_t smem_bytes = (size_t)P.smem_total_floats * 4;

    static bool configured = false;
    if (!configured) {
        TACO_CUDA(cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    TACO_CUDA(cudaMemsetAsync(a->workspace, 0, (size_t)P.ws.total * 8, st));
    void* args[] = {(void*)&P};
    TACO_CUDA(cudaLaunchCooperativeKernel((void*)decoder_kernel, dim3(NCTA), dim3(NTHR), args, smem_bytes, st));
    ++g_taco_launches;
    return 0;
}
