// decoder.cu -- the whole autoregressive attention decoder as ONE persistent dataflow kernel (design v4).
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
// History (C2: B=32, Tx=128, r=5; us per decoder step): v1 grid barrier + FFMA 58 -> v2 flag-in-data exchange 48
// -> v3 mma.sync 3xTF32, 12 dependent slots, weights partly streamed from L2 every step 24.8.  The v3 per-slot trace
// (profiles/r02_dec_trace_v3.log) showed ~3500 cycles per slot: ~2000 waiting for / ingesting K=512 operands through
// L2 (32 KB of {value,tag} words per CTA and slot = the chip's L2 bandwidth cap), ~500 epilogue, 300-1200 of weight
// streaming bookkeeping.  v4 attacks the length of the dependent chain and what sits on it:
//
//   * 9 dependent slots per step instead of 12 (weight-only algebra at pack time, exact up to re-association):
//       IN  G1 C1 G2 C2 G3 C3  OQP  A|P2
//     OQP computes, from s(t) alone, y(t) (output), q(t) = s.(W_out W_q) + b_out W_q and the FIRST PRE-NET LAYER OF
//     THE NEXT STEP p1(t+1) = relu(s.(W_out[:,last 80] W1) + b_out[last 80] W1 + b1) (free-running rows); A|P2 runs
//     the attention on warps 0-3 and the second pre-net layer on warps 4-7 concurrently; IN takes y(t-1) through
//     s(t-1).(W_out W_a[:80r] W_in[128:]) so y never travels.  Teacher-forced rows take p1 from a separate
//     off-chain stage PM (mel . W1), selected per row by the consumer.
//   * every K=512 contraction is split into the half that is on the dependent chain (x of the gates, r*h of the
//     candidate, ctx/p2 of the input projection) and the half whose operand has been known for a while (h(t-1), x,
//     s(t-1)).  The known half is multiplied in the shadow of the previous slot's L2 hop into the SAME accumulator
//     registers; only K=256 (16 KB of operand words per CTA) is ingested on the chain.
//   * ALL weights are resident: on-chain slices in shared memory (124 KB), off-chain slices in TENSOR MEMORY
//     (tcgen05.alloc of the SM's 256 KB TMEM, tcgen05.st once, tcgen05.ld.32x32b per use: every MMA weight fragment
//     is private to one lane, which is exactly TMEM's lane/column addressing).  Nothing is streamed per step.
//   * MMA orientation swapped: M = 16 weight columns of the CTA's slice, N = the 8 utterance rows of the CTA's row
//     group, so the 16-column gate stage has no zero-padded rows (half the MMAs of v3).
//   * attention: the four CTAs that share an utterance form a thread-block CLUSTER; partial softmax statistics and
//     partial contexts are exchanged through distributed shared memory ({value,tag} words pushed into the peer's
//     shared memory, polled locally), each CTA normalises and publishes 64 context columns and its own alignment
//     slice: the consumer ingests 256 context words per row instead of 4 x 256 partials + statistics.
//     tanh(k+q) = 1 - 2/(e^{2k} e^{2q} + 1) with e^{2k} precomputed once and e^{2q} once per step: one MUFU per
//     element instead of two.
//
// Unchanged from v3: grid = 128 co-resident CTAs x 256 threads, CTA (rg, cs) owns utterance rows 8rg..8rg+7 and the
// cs-th of 32 column slices of every dense stage; NO grid barrier: every exchanged activation is a 64-bit word
// {fp32 value, step tag} written with one st.b64 and read with polling 128-bit loads; 3xTF32 error-compensated
// tensor-core products (fp32-grade); GRU gate columns permuted so r,u of a hidden unit live in one CTA.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int NCTA = 128;
constexpr int NTHR = 512;      // 16 warps: 4 per SM sub-partition (the step is a chain of short dependent instruction sequences:
constexpr int NWARP = 16;      // with 2 warps per scheduler every fixed latency was exposed -- 'wait' was the top stall reason)
constexpr int RPG = 8;         // rows per row group
constexpr int NS = 32;         // column slices
constexpr int BPAD = 32;
constexpr int U = 256;         // decoder units
constexpr int AU = 256;        // attention units
constexpr int ENC = 256;       // memory depth
constexpr int MF = 80;
constexpr int KV_LD = 260;     // padded row stride of e^{2 keys} in smem
constexpr int PPITCH = 132;    // floats between the partial tiles of consecutive warps (128 + 4: cross-warp sums hit different banks)

// ---- weight segments (one K slab of one stage, per CTA column slice) ---------------------------------------------
enum SegId {
    SG_IN_S = 0, SG_IN_CP,
    SG_G_H0, SG_G_H1, SG_G_H2, SG_G_X0, SG_G_X1, SG_G_X2,
    SG_C_X0, SG_C_X1, SG_C_X2, SG_C_RH0, SG_C_RH1, SG_C_RH2,
    SG_Y, SG_QP, SG_PM, SG_P2, NSEG
};
// ---- exchange buffers: [4 row groups][nkt k-tiles][8 rows][8 words] of 64-bit {value, tag} words.  One k-tile of one
//      row group (512 B) is written by ONE producer CTA and read by one warp-wide 16-byte load: fully coalesced.
//      B_CP holds ctx (k-tiles 0..31) and p2 (k-tiles 32..47): the on-chain operand of the input projection.
enum BufId { B_S = 0, B_CP, B_Z, B_H0, B_H1, B_H2, B_RH0, B_RH1, B_RH2, B_Q, B_P1Y, B_P1M, NBUF };
__host__ __device__ inline int buf_nkt(int b) { return b == B_CP ? 48 : 32; }
// ---- stages (what "finish" does) --------------------------------------------------------------------------------
enum StageId { ST_IN = 0, ST_G0, ST_G1, ST_G2, ST_C0, ST_C1, ST_C2, ST_PM, ST_OQP, ST_AP2, ST_NONE };
// bias rows in smem ([row][16])
enum BiasRow { BR_IN = 0, BR_INX, BR_G0, BR_G1, BR_G2, BR_C0, BR_C1, BR_C2, BR_Y, BR_QP, BR_PM, BR_P2, NBR };
// item kinds: fragment width / tiles / k-tile slots per warp
enum ItemKind { IK_F2 = 0, IK_F4, IK_F4X2, IK_F2CP, IK_PM, IK_AP2 };   // 8 / 16 / 2x16 columns x 2 k-tile slots; [ctx|p2]: 3 slots
// output slots of the per-CTA output-offset table
enum OutId { O_Z = 0, O_RH0, O_RH1, O_RH2, O_H0, O_H1, O_H2, O_S, O_Q, O_P1Y, O_P1M, O_P2, NOUT };

struct SegDesc {
    int K;             // contraction length (multiple of 8)
    int FL;            // floats per lane and k-tile: 4 = 16 weight columns, 2 = 8 weight columns
    int nw;            // warps that split K (16, or 8 for the P2 stage)
    int kpw;           // k-tile slots per warp (2, 3 or 4; padded with zeros)
    int smem_off;      // float offset inside the resident smem region, or -1: lives in TMEM
    int tmem_col;      // column offset inside the warp's 128-column TMEM window
    int slice_floats;  // nw * kpw * 32 * FL
    int64_t g_off;     // float offset of slice 0 in the packed buffer
};

struct Item {          // one stage of the per-step program: [off-chain half] + on-chain half of one contraction, then finish
    int kind;          // ItemKind
    int stage;         // StageId finished after the on-chain half
    int pre_seg, pre_src, pre_tagd, pre_skip0;   // off-chain half: weight segment (-1: none), operand buffer, tag = t + tagd,
                                                 // 1 = operand does not exist at t == 0 (zero initial state)
    int on_seg, on_seg2, on_src, on_tagd, on_s0; // on-chain half; seg2 >= 0: second 16-column tile sharing the operand (OQP);
                                                 // on_s0 = leading k-tile slots that do not exist at t == 0
};

constexpr int MAX_ITEMS = 20;

struct DecParams {
    SegDesc seg[NSEG];
    Item items[MAX_ITEMS];
    int n_items;
    int64_t buf[NBUF];           // word offsets of the exchange buffers
    int64_t trace_off;
    taco_decoder_args a;
    const float *in_b, *gru_bg[3], *gru_bc[3], *out_b, *pre_b1, *pre_b2, *att_v;
    const float *b_qF, *b_p1F, *b_inS;     // fused biases (packed tail)
    int OUT, NCY, Tq;
    int smem_kv_off, smem_res_off, smem_total_floats;
};

// global column computed by weight-column m (0..15) of slice cs for a segment (-1 = padding)
__host__ __device__ inline int seg_col(int sg, int cs, int m, int NCY, int OUT) {
    if (sg >= SG_G_H0 && sg <= SG_G_X2) return (m < 8) ? (8 * cs + m) : (U + 8 * cs + (m - 8));     // r | u of units 8cs..8cs+7
    if (sg == SG_QP) return (m < 8) ? (8 * cs + m) : (AU + 8 * cs + (m - 8));                       // q | p1' (source = [W_qF | W_p1F])
    if (sg == SG_Y) { const int c = cs * NCY + m; return (m < NCY && c < OUT) ? c : -1; }
    if (sg == SG_P2) return (m < 4) ? (4 * cs + m) : -1;
    return (m < 8) ? (8 * cs + m) : -1;
}

// physical position of logical column k inside an exchange buffer row: within each group of 8,
// k and k+4 are adjacent so that one 16-byte load yields the (b0, b1) pair of an MMA B fragment.
__host__ __device__ inline int perm8(int k) {
    const int kk = k & 7;
    return (k & ~7) | ((kk < 4) ? 2 * kk : 2 * (kk - 4) + 1);
}

// ---------------------------------------------------------------------------------------------
// LL words: {value, tag}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint64_t* p, float v, uint32_t tag) {
    const uint64_t w = (uint64_t)__float_as_uint(v) | ((uint64_t)tag << 32);
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// polling load: ld.volatile measured 306 cycles per dependent L2 access on B200, ld.relaxed.gpu 467
// (scripts/ubench/latency.cu); both always observe L2.
__device__ __forceinline__ ulonglong2 ll_load2(const uint64_t* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool ll_ok(const ulonglong2& v, uint32_t tag) {
    return (((uint32_t)(v.x >> 32) ^ tag) | ((uint32_t)(v.y >> 32) ^ tag)) == 0u;
}
// Spin until both words carry `tag`; bounded so that a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void ll_spin(ulonglong2& v, const uint64_t* p, uint32_t tag) {
    uint32_t spins = 0;
    while (!ll_ok(v, tag)) {
        if (++spins > (1u << 24)) __trap();
        v = ll_load2(p);
    }
}
__device__ __forceinline__ float2 ll_wait2(const uint64_t* p, uint32_t tag) {
    ulonglong2 v = ll_load2(p);
    ll_spin(v, p, tag);
    return make_float2(__uint_as_float((uint32_t)v.x), __uint_as_float((uint32_t)v.y));
}
// the same protocol on words in (this CTA's) shared memory, written by cluster peers through DSMEM
__device__ __forceinline__ uint64_t sm_load(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ float sm_wait(const uint64_t* p, uint32_t tag) {
    uint64_t v = sm_load(p);
    uint32_t spins = 0;
    while ((uint32_t)(v >> 32) != tag) {
        if (++spins > (1u << 24)) __trap();
        v = sm_load(p);
    }
    return __uint_as_float((uint32_t)v);
}
// push a {value, tag} word into the shared memory of cluster CTA `rank` at the same offset as local `p`
__device__ __forceinline__ void dsm_store(const uint64_t* p, uint32_t rank, float v, uint32_t tag) {
    const uint64_t w = (uint64_t)__float_as_uint(v) | ((uint64_t)tag << 32);
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
    asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(ra), "l"(w) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float rcp_fast(float x) {       // one MUFU.RCP (<= 1 ulp); rcp(+inf) = 0
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// e^{2x} with the argument clamped so that products of two such factors stay finite: exact for |x| <= 20.8
// (tanh is saturated to the last bit far before that), graceful beyond
__device__ __forceinline__ float exp2x(float x) { return exp2f(fminf(fmaxf(x * 2.885390082f, -60.f), 60.f)); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---------------------------------------------------------------------------------------------
// tensor memory as lane-private weight storage
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tm_st2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_ld4(uint32_t taddr, float (&w)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, float (&w)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&w)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------
// 3xTF32 tensor-core contraction pieces
// ---------------------------------------------------------------------------------------------
// x = hi + lo with hi = the TF32-representable head (low 13 mantissa bits cleared) and lo = the exact
// remainder (the tensor core ignores lo's own low 13 bits: <= 2^-21 |x| dropped).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// D[16 weight columns x 8 utterance rows] += A[16x8] . B[8x8]
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// one k-tile: operand pair (x[row g][8kt+tg], x[row g][8kt+tg+4]) against one weight tile.
//   FL = 4: w = (W[k][m=g], W[k][m=g+8], W[k+4][m=g], W[k+4][m=g+8]);  FL = 2: w = (W[k][m=g], W[k+4][m=g]), rows 8..15 zero.
// The three 3xTF32 products go to three independent accumulators (no serialisation on the accumulator latency).
template <int FL>
__device__ __forceinline__ void ktile_mma(float (&acc)[3][4], uint32_t bh0, uint32_t bl0, uint32_t bh1, uint32_t bl1, const float* w) {
    uint32_t ah[4], al[4];
    if constexpr (FL == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(w[i], ah[i], al[i]);
    } else {
        split_tf32(w[0], ah[0], al[0]);
        split_tf32(w[1], ah[2], al[2]);
        ah[1] = al[1] = ah[3] = al[3] = 0u;
    }
    mma_tf32(acc[0], al[0], al[1], al[2], al[3], bh0, bh1);
    mma_tf32(acc[1], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
    mma_tf32(acc[2], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

// ---- operand ingest + multiply of one item -------------------------------------------------------------------------
// Slot i of warp-index `widx` (of `nw` warps splitting K) covers k-tile widx + nw*i; its 8 rows x 8 words sit at
// p0 + i*nw*64 (p0 already contains this lane's g*8 + 2*tg).  NS_ = slots per call (2, 3 or 4); s0 = first live slot.
template <int NS_>
__device__ __forceinline__ void issue_loads(ulonglong2 (&v)[4], const uint64_t* p0, int nkt, int widx, int nw, int s0) {
#pragma unroll
    for (int i = 0; i < NS_; ++i)
        if (i >= s0 && widx + nw * i < nkt) v[i] = ll_load2(p0 + (size_t)i * nw * 64);
}
// weight word: >= 0: float offset in the resident smem region; bit 31 set: TMEM column inside the warp's window.
// wslot = index of the first fragment slot used by this call in the segment's [warp][slot] block, tcoff = its TMEM column.
template <int FL, int MT, int NS_, bool TR>
__device__ __forceinline__ void consume(float (&acc)[2][3][4], ulonglong2 (&v)[4], const uint64_t* p0, int nkt, int widx, int nw, int s0,
                                        uint32_t tag, uint32_t w1, uint32_t w2, int wslot, int tcoff, const float* res_s, uint32_t tm_lane_col,
                                        int lane, long long* ck) {
    float w[MT][NS_ * FL];
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const uint32_t ww = mt ? w2 : w1;
        if (!(ww & 0x80000000u)) {
            const float* wp = res_s + ww + ((size_t)wslot * 32 + lane) * FL;
#pragma unroll
            for (int i = 0; i < NS_; ++i) {
                if constexpr (FL == 4) { const float4 t = *reinterpret_cast<const float4*>(wp + (size_t)i * 32 * FL); w[mt][4 * i] = t.x; w[mt][4 * i + 1] = t.y; w[mt][4 * i + 2] = t.z; w[mt][4 * i + 3] = t.w; }
                else                   { const float2 t = *reinterpret_cast<const float2*>(wp + (size_t)i * 32 * FL); w[mt][2 * i] = t.x; w[mt][2 * i + 1] = t.y; }
            }
        } else if constexpr (NS_ * FL == 4 || NS_ * FL == 8) {
            const uint32_t ta = tm_lane_col + (ww & 0x7fffffffu) + (uint32_t)tcoff;
            if constexpr (NS_ * FL == 8) tm_ld8(ta, w[mt]);
            else                         tm_ld4(ta, w[mt]);
        }
    }
    if (TR && ck) ck[1] = clock64();
#pragma unroll
    for (int i = 0; i < NS_; ++i) {
        if (i >= s0 && widx + nw * i < nkt) {          // warp-uniform
            ll_spin(v[i], p0 + (size_t)i * nw * 64, tag);
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(__uint_as_float((uint32_t)v[i].x), bh0, bl0);
            split_tf32(__uint_as_float((uint32_t)v[i].y), bh1, bl1);
            __syncwarp();                               // lanes arrive from divergent polling loops
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) ktile_mma<FL>(acc[mt], bh0, bl0, bh1, bl1, &w[mt][i * FL]);
        }
    }
    if (TR && ck) ck[2] = clock64();
}

struct ItemRec { uint32_t flags, pre_base; int pre_tagd; uint32_t pre_w; uint32_t on_base; int on_tagd; uint32_t on_w1, on_w2; int on_nkt, pad0, pad1, pad2; };   // 48 bytes, in smem
// flags: kind | stage << 8 | has_pre << 16 | pre_skip0 << 17 | on_s0 << 20

template <bool TR>     // TR: per-stage clock64 trace of CTA 0 (scripts/dec_time.py); the production instantiation has none of it
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NTHR, 1) decoder_kernel(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(16) float smem[];
    // smem map (floats): [0,16) tmem ptr | part 6400 | part2 512 | loc 384 | bias 256 | att 704 | xbuf 1024 | xstat 64 | itab 256 | kv | weights
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem);
    float* part_s = smem + 16;                                    // [3 buffers][16 warps][PPITCH]: two alternate, the third is OQP's 2nd tile
    float* part2_s = part_s + 6400;                               // [8 warps][64]   (P2, warps 8-15)
    float* loc_s = part2_s + 512;                                 // h_loc[3][64] | u_loc[64] | z_loc[64]
    float* bias_s = loc_s + 384;                                  // [NBR][16]
    float* att_s = bias_s + 256;                                  // eq[256] | v[256] | e[64] | p[64] | misc[64]
    uint64_t* xbuf = reinterpret_cast<uint64_t*>(att_s + 704);    // [2 parity][4 src][64] words
    uint64_t* xstat = xbuf + 512;                                 // [2 parity][4 src][2] words (+pad)
    ItemRec* itab = reinterpret_cast<ItemRec*>(att_s + 704 + 1024 + 64);   // [MAX_ITEMS] resolved stage records (20 x 12 words)
    uint32_t* otab = reinterpret_cast<uint32_t*>(itab + MAX_ITEMS);          // [NOUT] output word offsets of this CTA
    float* ek_s = smem + P.smem_kv_off;
    float* vals_s = ek_s + P.Tq * KV_LD;
    float* res_s = smem + P.smem_res_off;
    float* h_loc = loc_s;             // [3][8 cols][8 rows]
    float* u_loc = loc_s + 192;
    float* z_loc = loc_s + 256;
    float* eq_s = att_s;
    float* v_s = att_s + 256;
    float* e_s = att_s + 512;
    float* p_s = att_s + 576;
    float* misc_s = att_s + 640;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tg = lane & 3;            // MMA fragment coordinates
    const int cta = blockIdx.x;
    // dense stages: row group = the 32 consecutive CTAs that also run the attention of those 8 utterances, so the four
    // row groups are four INDEPENDENT 32-CTA machines on (mostly) neighbouring SMs: no exchange crosses a group
    const int rg = cta >> 5, cs = cta & 31;
    const int arow = cta >> 2;                          // attention: utterance (the cluster), quarter = cluster rank
    const uint32_t aq = cluster_rank();
    const taco_decoder_args& A = P.a;
    uint64_t* ws = reinterpret_cast<uint64_t*>(A.workspace);
    const int B = A.B, T = A.T, OUT = P.OUT, Tq = P.Tq, Tx = A.Tx, NCY = P.NCY;
    const int row0 = rg * RPG;
    const int myrow = row0 + g;                        // the utterance row this lane's B fragments belong to
    const int lane_w = g * 8 + 2 * tg;                  // this lane's word offset inside a k-tile block

    // ---- tensor memory: the whole 512-column TMEM of the SM (1 CTA/SM) ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 384; i += NTHR) loc_s[i] = 0.f;          // zero initial GRU states (cell.zero_state)
    for (int i = tid; i < 2 * (512 + 32); i += NTHR) reinterpret_cast<uint32_t*>(xbuf)[i] = 0u;   // tag 0 = invalid
    // biases of this CTA's columns -> smem
    if (tid < NBR * 16) {
        const int br = tid >> 4, m = tid & 15;
        float bv = 0.f;
        int col; const float* bp = nullptr;
        switch (br) {
            case BR_IN:  col = seg_col(SG_IN_CP, cs, m, NCY, OUT); bp = P.in_b; break;
            case BR_INX: col = seg_col(SG_IN_CP, cs, m, NCY, OUT); bp = P.b_inS; break;
            case BR_G0: case BR_G1: case BR_G2: col = seg_col(SG_G_X0, cs, m, NCY, OUT); bp = P.gru_bg[br - BR_G0]; break;
            case BR_C0: case BR_C1: case BR_C2: col = seg_col(SG_C_X0, cs, m, NCY, OUT); bp = P.gru_bc[br - BR_C0]; break;
            case BR_Y:   col = seg_col(SG_Y, cs, m, NCY, OUT); bp = P.out_b; break;
            case BR_QP:  col = (m < 8) ? 8 * cs + m : 8 * cs + (m - 8); bp = (m < 8) ? P.b_qF : P.b_p1F; break;
            case BR_PM:  col = seg_col(SG_PM, cs, m, NCY, OUT); bp = P.pre_b1; break;
            default:     col = seg_col(SG_P2, cs, m, NCY, OUT); bp = P.pre_b2; break;
        }
        if (bp && col >= 0) bv = __ldg(bp + col);
        bias_s[tid] = bv;
    }
    // resolved item records / output offsets (everything the per-item code would otherwise fetch through dependent
    // constant-bank loads: ~300 cycles per item in the first v4 trace)
    if (tid < P.n_items) {
        const Item& it = P.items[tid];
        ItemRec r;
        memset(&r, 0, sizeof(r));
        auto wword = [&](int sg) { const SegDesc& d = P.seg[sg]; return d.smem_off >= 0 ? (uint32_t)d.smem_off : (0x80000000u | (uint32_t)d.tmem_col); };
        r.flags = (uint32_t)it.kind | ((uint32_t)it.stage << 8) | ((it.pre_seg >= 0 ? 1u : 0u) << 16) | ((uint32_t)it.pre_skip0 << 17) | ((uint32_t)it.on_s0 << 20);
        if (it.pre_seg >= 0) {
            r.pre_base = (uint32_t)(P.buf[it.pre_src] + (int64_t)rg * buf_nkt(it.pre_src) * 64);
            r.pre_tagd = it.pre_tagd; r.pre_w = wword(it.pre_seg);
        }
        if (it.on_seg >= 0) {
            r.on_base = (uint32_t)(P.buf[it.on_src] + (int64_t)rg * buf_nkt(it.on_src) * 64);
            r.on_tagd = it.on_tagd; r.on_w1 = wword(it.on_seg); r.on_w2 = it.on_seg2 >= 0 ? wword(it.on_seg2) : 0u;
            r.on_nkt = P.seg[it.on_seg].K >> 3;
        }
        itab[tid] = r;
    }
    if (tid < NOUT) {
        const int bmap[NOUT] = {B_Z, B_RH0, B_RH1, B_RH2, B_H0, B_H1, B_H2, B_S, B_Q, B_P1Y, B_P1M, B_CP};
        const int b = bmap[tid];
        uint32_t o = (uint32_t)(P.buf[b] + (int64_t)rg * buf_nkt(b) * 64);
        if (tid == O_P2) o += (uint32_t)((32 + (cs >> 1)) * 64 + 0);     // p2 columns 4cs..4cs+3 live in k-tile 32 + cs/2 of B_CP
        else o += (uint32_t)(cs * 64);                                    // this CTA's 8 columns are k-tile cs
        otab[tid] = o;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // this warp's private TMEM window: lanes 32*(warp%4).., columns 128*(warp/4)..
    const uint32_t tm_lane_col = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u;

    // ---- one-time preload of the weight slices: smem segments by coalesced copies, TMEM segments by tcgen05.st ----
    for (int sg = 0; sg < NSEG; ++sg) {
        const SegDesc& d = P.seg[sg];
        const float* src = A.packed + d.g_off + (int64_t)cs * d.slice_floats;
        if (d.smem_off >= 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(res_s + d.smem_off);
            for (int i = tid; i < d.slice_floats / 4; i += NTHR) d4[i] = __ldg(s4 + i);
        } else {
            const int widx = (d.nw == NWARP) ? warp : warp - 8;
            if (widx >= 0) {                                            // warp-uniform
                for (int i = 0; i < d.kpw; ++i) {
                    const float* wp = src + ((size_t)(widx * d.kpw + i) * 32 + lane) * d.FL;
                    const uint32_t ta = tm_lane_col + (uint32_t)(d.tmem_col + i * d.FL);
                    if (d.FL == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(wp)); tm_st4(ta, t.x, t.y, t.z, t.w); }
                    else           { const float2 t = __ldg(reinterpret_cast<const float2*>(wp)); tm_st2(ta, t.x, t.y); }
                }
            }
        }
    }
    tm_wait_st();
    // e^{2 keys} / values slice of (arow, aq) -> smem (padded rows), zero for utterances >= B and positions >= Tx
    for (int i = tid; i < Tq * (ENC / 4); i += NTHR) {
        const int j = i / (ENC / 4), d4 = (i % (ENC / 4)) * 4;
        float4 kk = make_float4(0, 0, 0, 0), vv = kk;
        if (arow < B && (int)aq * Tq + j < Tx) {
            const int64_t gi = ((int64_t)arow * Tx + aq * Tq + j) * ENC + d4;
            kk = __ldg(reinterpret_cast<const float4*>(A.keys + gi));
            vv = __ldg(reinterpret_cast<const float4*>(A.values + gi));
        }
        kk.x = exp2x(kk.x); kk.y = exp2x(kk.y); kk.z = exp2x(kk.z); kk.w = exp2x(kk.w);   // tanh(k+q) = 1 - 2/(e^{2k} e^{2q} + 1)
        *reinterpret_cast<float4*>(ek_s + j * KV_LD + d4) = kk;
        *reinterpret_cast<float4*>(vals_s + j * ENC + d4) = vv;
    }
    for (int i = tid; i < AU; i += NTHR) v_s[i] = __ldg(P.att_v + i);
    const int my_len = (arow < B) ? A.text_length[arow] : 0;
    __syncthreads();
    if (warp == 0) {                                    // V0 = sum_d v_d  (e_j = V0 - 2 sum_d v_d / (E_jd + 1))
        float sv = 0.f;
        for (int i = lane; i < AU; i += 32) sv += v_s[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        if (lane == 0) misc_s[0] = sv;
    }
    __syncthreads();
    cluster_sync_all();                                 // peers' shared memory is initialised before anyone pushes into it

    float acc[2][3][4];
#pragma unroll
    for (int a0 = 0; a0 < 2; ++a0)
#pragma unroll
        for (int a1 = 0; a1 < 3; ++a1)
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2) acc[a0][a1][a2] = 0.f;
    ulonglong2 v[4];                                    // operand words in flight

    int par = 0;                                        // parity of the partial-tile buffer
    const bool tracer = (cta == 0 && tid == 0 && A.step_ns != nullptr);
    int64_t trace_i = 0;
    long long* cktr = nullptr;                          // per-item clock64 trace (CTA 0, thread 0, steps 10..13)

    // =========================================================================================================
    // finish a stage: per-warp partial tiles -> smem, cross-warp sum, epilogue (one thread per output)
    // =========================================================================================================
    auto store_partials = [&](float* part, float* part_b, int MT, bool full16) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (mt < MT) {
                float* pw = (mt ? part_b : part) + warp * PPITCH;
                *reinterpret_cast<float2*>(pw + g * 8 + 2 * tg) =
                    make_float2((acc[mt][0][0] + acc[mt][1][0]) + acc[mt][2][0], (acc[mt][0][1] + acc[mt][1][1]) + acc[mt][2][1]);
                if (full16)
                    *reinterpret_cast<float2*>(pw + (g + 8) * 8 + 2 * tg) =
                        make_float2((acc[mt][0][2] + acc[mt][1][2]) + acc[mt][2][2], (acc[mt][0][3] + acc[mt][1][3]) + acc[mt][2][3]);
            }
        }
#pragma unroll
        for (int a0 = 0; a0 < 2; ++a0)
#pragma unroll
            for (int a1 = 0; a1 < 3; ++a1)
#pragma unroll
                for (int a2 = 0; a2 < 4; ++a2) acc[a0][a1][a2] = 0.f;
    };
    // cross-warp sum of one output by one thread.  (Spreading each sum over 2-4 lanes + shuffles so that all 16 warps share
    // the epilogue was tried: 16.0 instead of 15.3 us/step -- the extra shuffles cost more than the shorter chains save.)
    auto sum16 = [&](const float* part, int pidx) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; w += 2) { s0 += part[w * PPITCH + pidx]; s1 += part[(w + 1) * PPITCH + pidx]; }
        return s0 + s1;
    };
    // epilogue thread mapping: lanes run over the 8 weight columns first (8-byte words of one row are then
    // contiguous: a warp writes two full 128-byte lines), rows next, the second 8 columns (G: u gates) last.
    const int e_brow = (tid >> 3) & 7, e_wc = tid & 7, e_hi = (tid >> 6) & 1;
    const int e_pidx = (e_wc + 8 * e_hi) * 8 + e_brow;       // index into a [16 cols][8 rows] partial tile
    const int e_loc = e_wc * 8 + e_brow;                     // index into the [8 cols][8 rows] local state arrays
    const int e_row = row0 + e_brow;
    const int e_word = e_brow * 8 + perm8(e_wc);             // word inside this CTA's k-tile block of an exchange buffer

    // ---- attention of step t on warps 0..7 (256 threads): scores, partial softmax / context, cluster merge ----
    auto attention = [&](int t) {
        const uint32_t tag = (uint32_t)t + 1;
        const int xp = t & 1;
        uint64_t* ctx_out = ws + P.buf[B_CP] + (int64_t)(arow >> 3) * 48 * 64 + (8 * (int)aq + ((tid & 63) >> 3)) * 64 + (arow & 7) * 8 + perm8(tid & 7);
        if (arow >= B) {                                // padding utterance: publish a zero context (consumers poll every row)
            if (tid < 64) ll_store(ctx_out, 0.f, tag + 1);
            return;
        }
        if (tid < 128) {   // q(t) of this utterance -> e^{2q}
            const float2 qq = ll_wait2(ws + P.buf[B_Q] + (int64_t)(arow >> 3) * 2048 + (tid >> 2) * 64 + (arow & 7) * 8 + 2 * (tid & 3), tag);
            const int k0 = ((tid >> 2) << 3) + (tid & 3);        // physical pair (2p, 2p+1) = logical k0, k0+4
            eq_s[k0] = exp2x(qq.x);
            eq_s[k0 + 4] = exp2x(qq.y);
        }
        if (TR && cktr) cktr[0] = clock64();
        named_bar(1, 256);
        if (TR && cktr) cktr[1] = clock64();
        {   // scores: lane owns 8 consecutive depth indices (its e^{2q}, v slices stay in registers), warp w owns
            // positions j = w, w+8, ...; four positions are reduced over the warp with one transposing butterfly
            float eq[8], vv[8];
            {
                const float4 q0 = *reinterpret_cast<const float4*>(eq_s + 8 * lane), q1 = *reinterpret_cast<const float4*>(eq_s + 8 * lane + 4);
                const float4 v0 = *reinterpret_cast<const float4*>(v_s + 8 * lane), v1 = *reinterpret_cast<const float4*>(v_s + 8 * lane + 4);
                eq[0] = q0.x; eq[1] = q0.y; eq[2] = q0.z; eq[3] = q0.w; eq[4] = q1.x; eq[5] = q1.y; eq[6] = q1.z; eq[7] = q1.w;
                vv[0] = v0.x; vv[1] = v0.y; vv[2] = v0.z; vv[3] = v0.w; vv[4] = v1.x; vv[5] = v1.y; vv[6] = v1.z; vv[7] = v1.w;
            }
            const float V0 = misc_s[0];
            for (int jb = 0; jb < Tq; jb += 32) {       // 4 positions per warp and round
                float a4[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = jb + warp + 8 * jj;
                    float a = 0.f;
                    if (j < Tq) {                       // warp-uniform
                        const float4 k0 = *reinterpret_cast<const float4*>(ek_s + j * KV_LD + 8 * lane);
                        const float4 k1 = *reinterpret_cast<const float4*>(ek_s + j * KV_LD + 8 * lane + 4);
                        a = fmaf(vv[0], rcp_fast(fmaf(k0.x, eq[0], 1.0f)), a);
                        a = fmaf(vv[1], rcp_fast(fmaf(k0.y, eq[1], 1.0f)), a);
                        a = fmaf(vv[2], rcp_fast(fmaf(k0.z, eq[2], 1.0f)), a);
                        a = fmaf(vv[3], rcp_fast(fmaf(k0.w, eq[3], 1.0f)), a);
                        a = fmaf(vv[4], rcp_fast(fmaf(k1.x, eq[4], 1.0f)), a);
                        a = fmaf(vv[5], rcp_fast(fmaf(k1.y, eq[5], 1.0f)), a);
                        a = fmaf(vv[6], rcp_fast(fmaf(k1.z, eq[6], 1.0f)), a);
                        a = fmaf(vv[7], rcp_fast(fmaf(k1.w, eq[7], 1.0f)), a);
                    }
                    a4[jj] = a;
                }
                // transposing butterfly: after the two halving steps lane group (lane>>3) holds position jj = lane>>3
                float b2[2], b1;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = (lane & 16) ? a4[k] : a4[k + 2];
                    const float keep = (lane & 16) ? a4[k + 2] : a4[k];
                    b2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                {
                    const float send = (lane & 8) ? b2[0] : b2[1];
                    const float keep = (lane & 8) ? b2[1] : b2[0];
                    b1 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                b1 += __shfl_xor_sync(0xffffffffu, b1, 4);
                b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
                b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
                const int jj = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
                const int j = jb + warp + 8 * jj;
                if ((lane & 7) == 0 && j < Tq) e_s[j] = ((int)aq * Tq + j < my_len) ? fmaf(-2.0f, b1, V0) : -INFINITY;
            }
        }
        if (TR && cktr) cktr[2] = clock64();
        named_bar(1, 256);
        if (TR && cktr) cktr[4] = clock64();
        // quarter statistics (every warp redundantly; Tq <= 64)
        const float e0 = (lane < Tq) ? e_s[lane] : -INFINITY;
        const float e1 = (lane + 32 < Tq) ? e_s[lane + 32] : -INFINITY;
        float m = fmaxf(e0, e1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float p0 = (e0 == -INFINITY) ? 0.f : __expf(e0 - m);
        const float p1 = (e1 == -INFINITY) ? 0.f : __expf(e1 - m);
        float ssum = p0 + p1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        if (warp == 0) {
            if (lane < Tq) p_s[lane] = p0;
            if (lane + 32 < Tq) p_s[lane + 32] = p1;
        }
        if (tid < 4) {                                   // statistics -> every CTA of the cluster (incl. this one)
            dsm_store(xstat + (xp * 4 + aq) * 2, (uint32_t)tid, m, tag);
            dsm_store(xstat + (xp * 4 + aq) * 2 + 1, (uint32_t)tid, ssum, tag);
        }
        named_bar(1, 256);
        if (TR && cktr) cktr[5] = clock64();
        {   // partial context: column tid; pushed to the CTA that owns it (64 columns per CTA)
            float c0 = 0.f, c1 = 0.f;
#pragma unroll 4
            for (int j = 0; j + 1 < Tq; j += 2) {
                c0 = fmaf(p_s[j], vals_s[j * ENC + tid], c0);
                c1 = fmaf(p_s[j + 1], vals_s[(j + 1) * ENC + tid], c1);
            }
            if (Tq & 1) c0 = fmaf(p_s[Tq - 1], vals_s[(Tq - 1) * ENC + tid], c0);
            dsm_store(xbuf + (xp * 4 + aq) * 64 + (tid & 63), (uint32_t)tid >> 6, c0 + c1, tag);
        }
        if (TR && cktr) cktr[6] = clock64();
        if (tid >= 128) return;
        // merge: global max / sum from the four quarter statistics
        float mq[4], wq[4], M = -INFINITY, S = 0.f, w_own = 0.f;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) { mq[qd] = sm_wait(xstat + (xp * 4 + qd) * 2, tag); M = fmaxf(M, mq[qd]); }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float sq = sm_wait(xstat + (xp * 4 + qd) * 2 + 1, tag);
            wq[qd] = (mq[qd] == -INFINITY) ? 0.f : __expf(mq[qd] - M);
            S = fmaf(wq[qd], sq, S);
            if (qd == (int)aq) w_own = wq[qd];
        }
        const float inv = (S > 0.f) ? 1.0f / S : 0.f;
        if (tid < 64) {                                  // final context column 64 aq + tid  (tag of the step that consumes it)
            float c = 0.f;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) c = fmaf(wq[qd], sm_wait(xbuf + (xp * 4 + qd) * 64 + tid, tag), c);
            ll_store(ctx_out, c * inv, tag + 1);
            if (TR && cktr) cktr[7] = clock64();
        } else {                                         // alignments of this quarter (AttentionWrapper alignment_history)
            const int j = tid - 64;
            if (j < Tq && (int)aq * Tq + j < Tx) A.align[((int64_t)arow * T + t) * Tx + aq * Tq + j] = p_s[j] * (w_own * inv);
        }
    };

    // ---- second pre-net layer of step tb on warps 8..15: p2(tb) = relu(p1(tb) . W2 + b2) ----
    auto prenet2 = [&](int tb) {
        const uint32_t tag = (uint32_t)tb + 1;
        const int widx = warp - 8;
        bool from_y = (tb > 0);
        if (A.mode == TACO_DEC_TEACHER) from_y = false;
        else if (A.mode == TACO_DEC_SCHED) from_y = (tb > 0) && (myrow < B) && (A.sample_mask[(int64_t)(tb - 1) * B + myrow] != 0);
        const uint64_t* p0 = ws + P.buf[from_y ? B_P1Y : B_P1M] + (int64_t)rg * 2048 + widx * 64 + lane_w;
        const SegDesc& sd = P.seg[SG_P2];
        const uint32_t w1 = sd.smem_off >= 0 ? (uint32_t)sd.smem_off : (0x80000000u | (uint32_t)sd.tmem_col);
        issue_loads<4>(v, p0, 32, widx, 8, 0);
        consume<2, 1, 4, false>(acc, v, p0, 32, widx, 8, 0, tag, w1, w1, widx * 4, 0, res_s, tm_lane_col, lane, nullptr);
        float* pw = part2_s + widx * 64;
        *reinterpret_cast<float2*>(pw + g * 8 + 2 * tg) =
            make_float2((acc[0][0][0] + acc[0][1][0]) + acc[0][2][0], (acc[0][0][1] + acc[0][1][1]) + acc[0][2][1]);
#pragma unroll
        for (int a1 = 0; a1 < 3; ++a1)
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2) acc[0][a1][a2] = 0.f;
        named_bar(2, 256);
        const int e = tid - 256;
        if (e < 32) {                                    // 8 rows x 4 columns, columns fastest
            const int wcol = e & 3, brow = e >> 2, row = row0 + brow, col = 4 * cs + wcol, pi = wcol * 8 + brow;
            float pv = bias_s[BR_P2 * 16 + wcol];
#pragma unroll
            for (int w = 0; w < 8; ++w) pv += part2_s[w * 64 + pi];
            pv = fmaxf(pv, 0.f);
            if (A.keep2 && row < B && tb < T) pv = A.keep2[((int64_t)tb * B + row) * 128 + col] ? pv * A.keep_scale : 0.f;
            ll_store(ws + otab[O_P2] + brow * 8 + perm8(4 * (cs & 1) + wcol), pv, tag);
        }
        named_bar(2, 256);                               // part2_s may be rewritten
    };

    // ---- first pre-net layer of step tb from the teacher frame (or zeros): p1m(tb) = relu(x . W1 + b1) ----
    auto prenet1_teacher = [&](int tb) {
        const bool have = (A.mode != TACO_DEC_INFER) && (myrow < B) && (tb < T);
        const SegDesc& sd = P.seg[SG_PM];
        float w[4];
        if (sd.smem_off >= 0) {
            const float2 t2 = *reinterpret_cast<const float2*>(res_s + sd.smem_off + ((size_t)(warp * sd.kpw) * 32 + lane) * 2);
            w[0] = t2.x; w[1] = t2.y; w[2] = w[3] = 0.f;
        } else {
            __syncwarp();
            tm_ld4(tm_lane_col + (uint32_t)sd.tmem_col, w);
        }
        if (warp < MF / 8) {                             // k-tile = warp (10 k-tiles); warp-uniform
            float xa = 0.f, xb = 0.f;
            if (have) {
                const float* mp = A.mel + ((int64_t)myrow * T + tb) * OUT + (OUT - MF) + warp * 8 + tg;
                xa = __ldg(mp); xb = __ldg(mp + 4);
            }
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(xa, bh0, bl0);
            split_tf32(xb, bh1, bl1);
            __syncwarp();
            ktile_mma<2>(acc[0], bh0, bl0, bh1, bl1, &w[0]);
        }
    };

    auto finish = [&](int stage, int t) {
        float* part = part_s + par * (NWARP * PPITCH);
        float* part_b = part_s + 2 * (NWARP * PPITCH);               // second tile of OQP (its previous readers are a full step behind)
        par ^= 1;
        const bool full16 = (stage >= ST_G0 && stage <= ST_G2) || stage == ST_OQP;
        store_partials(part, part_b, stage == ST_OQP ? 2 : 1, full16);
        if (TR && cktr) cktr[4] = clock64();
        __syncthreads();
        if (TR && cktr) cktr[5] = clock64();
        const uint32_t tag = (uint32_t)t + 1;
        switch (stage) {
            case ST_IN:
                if (tid < 64) {
                    float zv = sum16(part, e_pidx) + bias_s[BR_IN * 16 + e_wc];
                    if (t > 0) zv += bias_s[BR_INX * 16 + e_wc];
                    z_loc[e_loc] = zv;
                    ll_store(ws + otab[O_Z] + e_word, zv, tag);
                }
                break;
            case ST_G0: case ST_G1: case ST_G2:
                if (tid < 128) {
                    const int gi = stage - ST_G0;
                    const float gt = sigmoidf_acc(sum16(part, e_pidx) + bias_s[(BR_G0 + gi) * 16 + e_wc + 8 * e_hi]);
                    if (!e_hi) ll_store(ws + otab[O_RH0 + gi] + e_word, gt * h_loc[gi * 64 + e_loc], tag);   // r * h
                    else u_loc[e_loc] = gt;                                                                // u stays local
                }
                break;
            case ST_C0: case ST_C1: case ST_C2:
                if (tid < 64) {
                    const int gi = stage - ST_C0;
                    const float cnd = tanhf_acc(sum16(part, e_pidx) + bias_s[(BR_C0 + gi) * 16 + e_wc]);
                    const float uu = u_loc[e_loc];
                    const float hn = uu * h_loc[gi * 64 + e_loc] + (1.0f - uu) * cnd;
                    h_loc[gi * 64 + e_loc] = hn;
                    ll_store(ws + otab[O_H0 + gi] + e_word, hn, tag);
                    if (gi == 2) ll_store(ws + otab[O_S] + e_word, z_loc[e_loc] + hn, tag);
                    if (A.h_save && e_row < B) A.h_save[(((int64_t)gi * T + t) * B + e_row) * U + 8 * cs + e_wc] = hn;   // training: BPTT input
                }
                break;
            case ST_PM:
                if (tid < 64) {                              // here t = tb, the step the pre-net output belongs to
                    const int col = 8 * cs + e_wc;
                    float pv = fmaxf(sum16(part, e_pidx) + bias_s[BR_PM * 16 + e_wc], 0.f);
                    if (A.keep1 && e_row < B && t < T) pv = A.keep1[((int64_t)t * B + e_row) * 256 + col] ? pv * A.keep_scale : 0.f;
                    ll_store(ws + otab[O_P1M] + e_word, pv, tag);
                }
                break;
            case ST_OQP: if (tid < 256) {
                const int mt = tid >> 7;
                const float v0 = sum16(mt ? part_b : part, e_pidx) + bias_s[(BR_Y + mt) * 16 + e_wc + 8 * e_hi];
                if (mt == 0) {                               // y(t): the output itself (nothing downstream reads it)
                    const int wcol = e_wc + 8 * e_hi, col = cs * NCY + wcol;
                    if (wcol < NCY && col < OUT && e_row < B) A.y[((int64_t)e_row * T + t) * OUT + col] = v0;
                } else if (!e_hi) {                          // q(t)
                    ll_store(ws + otab[O_Q] + e_word, v0, tag);
                } else {                                     // p1(t+1) of free-running rows
                    const int col = 8 * cs + e_wc, tb = t + 1;
                    float pv = fmaxf(v0, 0.f);
                    if (A.keep1 && e_row < B && tb < T) pv = A.keep1[((int64_t)tb * B + e_row) * 256 + col] ? pv * A.keep_scale : 0.f;
                    ll_store(ws + otab[O_P1Y] + e_word, pv, tag + 1);
                }
            } break;
            default: break;
        }
        if (TR && cktr) cktr[6] = clock64();
        if (TR && tracer) ws[P.trace_off + (trace_i++)] = globaltimer_ns();
    };

    // =========================================================================================================
    // prologue: p1(0) (teacher frame 0 or zeros), p2(0)
    // =========================================================================================================
    prenet1_teacher(0);
    finish(ST_PM, 0);
    if (warp >= 8) prenet2(0);

    // =========================================================================================================
    // T decoder steps x the item program
    // =========================================================================================================
    const int n_items = P.n_items;
    const uint32_t itab_a = smem_u32(itab);
    ulonglong2 vp[4];                                   // operand words of the off-chain half (prefetched one finish() early)
    bool pref = false;                                  // vp[] already holds the loads of the current stage's off-chain half
    auto ld_rec = [&](int ii, uint4& r0, uint4& r1, uint32_t& nkt) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(itab_a + ii * 48));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(itab_a + ii * 48 + 16));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nkt) : "r"(itab_a + ii * 48 + 32));
    };
    uint4 nr0, nr1; uint32_t nnkt;
    ld_rec(0, nr0, nr1, nnkt);
    for (int t = 0; t < T; ++t) {
        if (tracer) A.step_ns[t] = globaltimer_ns();
        for (int ii = 0; ii < n_items; ++ii) {
            // r0 = flags, pre_base, pre_tagd, pre_w;  r1 = on_base, on_tagd, on_w1, on_w2 -- fetched one stage ahead
            const uint4 r0 = nr0, r1 = nr1; const uint32_t on_nkt = nnkt;
            ld_rec(ii + 1 < n_items ? ii + 1 : 0, nr0, nr1, nnkt);
            const int kind = r0.x & 0xff, stage = (r0.x >> 8) & 0xff;
            long long* ck = nullptr;
            if (TR && tracer && t >= 10 && t < 14) { ck = reinterpret_cast<long long*>(ws + P.trace_off + 16 * T + 16 + ((t - 10) * MAX_ITEMS + ii) * 8); ck[3] = clock64(); }
            cktr = ck;
            if (kind == IK_AP2) {
                if (warp < 8) attention(t);
                else if (t + 1 < T) prenet2(t + 1);
                if (TR && tracer) ws[P.trace_off + (trace_i++)] = globaltimer_ns();
                continue;
            }
            if (kind == IK_PM) {
                if (t + 1 < T) { prenet1_teacher(t + 1); finish(ST_PM, t + 1); }
                continue;
            }
            // ---- on-chain operand: loads in flight first, they travel while the off-chain half is multiplied ----
            const uint64_t* pon = ws + r1.x + warp * 64 + lane_w;
            const int s0 = (t == 0) ? (int)((r0.x >> 20) & 0xf) : 0;
            if (kind == IK_F2CP) issue_loads<3>(v, pon, (int)on_nkt, warp, NWARP, s0);
            else                 issue_loads<2>(v, pon, (int)on_nkt, warp, NWARP, 0);
            if (TR && ck) ck[0] = clock64();
            const bool has_pre = ((r0.x >> 16) & 1) && !(t == 0 && ((r0.x >> 17) & 1));
            const uint64_t* ppre = ws + r0.y + warp * 64 + lane_w;
            const uint32_t ptag = (uint32_t)(t + (int)r0.z), otag = (uint32_t)(t + (int)r1.y);
            if (has_pre && !pref) issue_loads<2>(vp, ppre, 32, warp, NWARP, 0);
            if (kind == IK_F4) {
                if (has_pre) consume<4, 1, 2, false>(acc, vp, ppre, 32, warp, NWARP, 0, ptag, r0.w, 0u, warp * 2, 0, res_s, tm_lane_col, lane, nullptr);
                consume<4, 1, 2, TR>(acc, v, pon, 32, warp, NWARP, 0, otag, r1.z, 0u, warp * 2, 0, res_s, tm_lane_col, lane, ck);
            } else if (kind == IK_F2) {
                if (has_pre) consume<2, 1, 2, false>(acc, vp, ppre, 32, warp, NWARP, 0, ptag, r0.w, 0u, warp * 2, 0, res_s, tm_lane_col, lane, nullptr);
                consume<2, 1, 2, TR>(acc, v, pon, 32, warp, NWARP, 0, otag, r1.z, 0u, warp * 2, 0, res_s, tm_lane_col, lane, ck);
            } else if (kind == IK_F2CP) {
                if (has_pre) consume<2, 1, 2, false>(acc, vp, ppre, 32, warp, NWARP, 0, ptag, r0.w, 0u, warp * 2, 0, res_s, tm_lane_col, lane, nullptr);
                consume<2, 1, 3, TR>(acc, v, pon, 48, warp, NWARP, s0, otag, r1.z, 0u, warp * 3, 0, res_s, tm_lane_col, lane, ck);
            } else {
                consume<4, 2, 2, TR>(acc, v, pon, 32, warp, NWARP, 0, otag, r1.z, r1.w, warp * 2, 0, res_s, tm_lane_col, lane, ck);
            }
            pref = false;
            // the next stage's off-chain operand is already known: put its loads in flight now, they travel while this
            // stage's epilogue runs
            if (ii + 1 < n_items && ((nr0.x >> 16) & 1)) {
                issue_loads<2>(vp, ws + nr0.y + warp * 64 + lane_w, 32, warp, NWARP, 0);
                pref = true;
            }
            finish(stage, t);
        }
    }

    // ---- teardown: nobody may leave while a peer can still push into its shared memory; free the TMEM ----
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- packing: TF [K][N] -> per-slice MMA A fragments [cs][warp][k-tile slot][lane][FL] ----
__global__ void pack_seg_kernel(const float* __restrict__ W, int ldw, int k0, int K, int FL, int nw, int kpw, int sg, int NCY, int OUT,
                                float* __restrict__ dst) {
    const int slice = nw * kpw * 32 * FL;
    const int64_t total = (int64_t)NS * slice;
    const int nkt = K / 8;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int cs = (int)(idx / slice);
        int r = (int)(idx % slice);
        const int e = r % FL; r /= FL;
        const int lane = r & 31; r >>= 5;
        const int i = r % kpw;
        const int widx = r / kpw;
        const int kt = widx + nw * i;
        const int g = lane >> 2, tg = lane & 3;
        // FL = 4: e = (k, m=g), (k, m=g+8), (k+4, m=g), (k+4, m=g+8);  FL = 2: e = (k, m=g), (k+4, m=g)
        const int m = (FL == 4) ? g + 8 * (e & 1) : g;
        const int k = kt * 8 + tg + 4 * ((FL == 4) ? (e >> 1) : e);
        float v = 0.f;
        if (kt < nkt) {
            const int col = seg_col(sg, cs, m, NCY, OUT);
            if (col >= 0) v = W[(int64_t)(k0 + k) * ldw + col];
        }
        dst[idx] = v;
    }
}

// weight-only precompute (double accumulation): C = A . B (+ addv)
__global__ void matmul_naive_kernel(const float* __restrict__ Amat, int lda, const float* __restrict__ Bmat, int ldb,
                                    float* __restrict__ Cmat, int ldc, int M, int N, int K, const float* __restrict__ addv) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i % N);
        double a = addv ? (double)addv[n] : 0.0;
        for (int k = 0; k < K; ++k) a += (double)Amat[(int64_t)m * lda + k] * (double)Bmat[(int64_t)k * ldb + n];
        Cmat[(int64_t)m * ldc + n] = (float)a;
    }
}

// ---- fused linear stages (weight-only algebra; sums are re-associated, ~1e-6 relative) --------------------------
//   W_cpF  = [ W_a[80r:80r+256] . W_in[128:384] ; W_in[0:128] ]    [ctx(t-1) | p2(t)] -> z(t)
//   M1     = W_a[0:80r] . W_in[128:384];  W_sF = W_out . M1;  b_inS = b_out . M1        y(t-1) = s(t-1).W_out + b_out -> z(t)
//   W_qp   = [ W_out . W_q | W_out[:, last 80] . W1 ];  b_qF = b_out . W_q;  b_p1F = b_out[last 80] . W1 + b1
struct FusedTail { int64_t w_cpF, m1, w_sF, w_qp, b_qF, b_p1F, b_inS, total; };
FusedTail fused_tail(int r, int64_t slices_total) {
    const int OUT = MF * r;
    FusedTail f;
    int64_t o = (slices_total + 63) / 64 * 64;
    f.w_cpF = o;  o += (int64_t)(ENC + 128) * U;          // [W_ctxF ; W_in[0:128]]: the on-chain operand [ctx | p2]
    f.m1 = o;     o += (int64_t)OUT * U;
    f.w_sF = o;   o += (int64_t)U * U;
    f.w_qp = o;   o += (int64_t)U * 2 * AU;
    f.b_qF = o;   o += AU;
    f.b_p1F = o;  o += 256;
    f.b_inS = o;  o += U;
    f.total = o;
    return f;
}

void build_seg_table(SegDesc* sd, int64_t* total_floats) {
    auto set = [&](int id, int K, int FL, int nw, int kpw) { sd[id].K = K; sd[id].FL = FL; sd[id].nw = nw; sd[id].kpw = kpw; };
    set(SG_IN_S, 256, 2, 16, 2); set(SG_IN_CP, 384, 2, 16, 3);
    for (int i = 0; i < 3; ++i) {
        set(SG_G_H0 + i, 256, 4, 16, 2); set(SG_G_X0 + i, 256, 4, 16, 2);
        set(SG_C_X0 + i, 256, 2, 16, 2); set(SG_C_RH0 + i, 256, 2, 16, 2);
    }
    set(SG_Y, 256, 4, 16, 2); set(SG_QP, 256, 4, 16, 2); set(SG_PM, MF, 2, 16, 2); set(SG_P2, 256, 2, 8, 4);
    int64_t off = 0;
    for (int s = 0; s < NSEG; ++s) {
        sd[s].slice_floats = sd[s].nw * sd[s].kpw * 32 * sd[s].FL;
        sd[s].g_off = off; sd[s].smem_off = -1; sd[s].tmem_col = 0;
        off += (int64_t)NS * sd[s].slice_floats;
    }
    *total_floats = off;
}

void build_ws_layout(int64_t* buf, int64_t* total) {
    int64_t o = 0;
    for (int b = 0; b < NBUF; ++b) { buf[b] = o; o += (int64_t)4 * buf_nkt(b) * 64; }
    *total = o;
}

inline int trace_words(int T) { return 16 * (T > 0 ? T : 0) + 16 + 4 * MAX_ITEMS * 8 + 16; }

}  // namespace

extern "C" size_t taco_decoder_packed_bytes(int r) {
    SegDesc sd[NSEG]; int64_t tot;
    build_seg_table(sd, &tot);
    return (size_t)fused_tail(r, tot).total * 4;
}

extern "C" size_t taco_decoder_workspace_bytes(int B, int Tx, int T, int r) {
    (void)B; (void)Tx; (void)r;
    int64_t buf[NBUF], total;
    build_ws_layout(buf, &total);
    return (size_t)(total + trace_words(T)) * 8;
}

extern "C" int taco_decoder_pack(const taco_decoder_weights* w, int r, float* packed, void* stream) {
    TACO_CHECK(w && packed, "taco_decoder_pack: NULL");
    TACO_CHECK(r >= 1 && MF * r <= 512, "taco_decoder_pack: r=%d out of range (80r <= 512)", r);
    TACO_CHECK(w->att_Wa && w->in_W && w->out_W && w->att_Wq && w->out_b && w->pre_W1 && w->pre_b1 && w->pre_W2, "taco_decoder_pack: NULL weight");
    for (int i = 0; i < 3; ++i) TACO_CHECK(w->gru_Wg[i] && w->gru_Wc[i], "taco_decoder_pack: NULL GRU weight");
    SegDesc sd[NSEG]; int64_t tot;
    build_seg_table(sd, &tot);
    cudaStream_t stm = (cudaStream_t)stream;
    const int OUT = MF * r;
    const int NCY = (OUT + NS - 1) / NS;
    const FusedTail ft = fused_tail(r, tot);
    float *w_cpF = packed + ft.w_cpF, *m1 = packed + ft.m1, *w_sF = packed + ft.w_sF, *w_qp = packed + ft.w_qp;
    const float* w_inA = w->in_W + (int64_t)128 * U;           // W_in[128:384]
    matmul_naive_kernel<<<256, 256, 0, stm>>>(w->att_Wa + (int64_t)OUT * AU, AU, w_inA, U, w_cpF, U, ENC, U, AU, nullptr);
    TACO_LAUNCH_CHECK();
    TACO_CUDA(cudaMemcpyAsync(w_cpF + (int64_t)ENC * U, w->in_W, (size_t)128 * U * sizeof(float), cudaMemcpyDeviceToDevice, stm));
    matmul_naive_kernel<<<400, 256, 0, stm>>>(w->att_Wa, AU, w_inA, U, m1, U, OUT, U, AU, nullptr);
    TACO_LAUNCH_CHECK();
    matmul_naive_kernel<<<256, 256, 0, stm>>>(w->out_W, OUT, m1, U, w_sF, U, U, U, OUT, nullptr);
    TACO_LAUNCH_CHECK();
    matmul_naive_kernel<<<1, 256, 0, stm>>>(w->out_b, OUT, m1, U, packed + ft.b_inS, U, 1, U, OUT, nullptr);
    TACO_LAUNCH_CHECK();
    matmul_naive_kernel<<<256, 256, 0, stm>>>(w->out_W, OUT, w->att_Wq, AU, w_qp, 2 * AU, U, AU, OUT, nullptr);
    TACO_LAUNCH_CHECK();
    matmul_naive_kernel<<<256, 256, 0, stm>>>(w->out_W + (OUT - MF), OUT, w->pre_W1, 256, w_qp + AU, 2 * AU, U, 256, MF, nullptr);
    TACO_LAUNCH_CHECK();
    matmul_naive_kernel<<<1, 256, 0, stm>>>(w->out_b, OUT, w->att_Wq, AU, packed + ft.b_qF, AU, 1, AU, OUT, nullptr);
    TACO_LAUNCH_CHECK();
    matmul_naive_kernel<<<1, 256, 0, stm>>>(w->out_b + (OUT - MF), MF, w->pre_W1, 256, packed + ft.b_p1F, 256, 1, 256, MF, w->pre_b1);
    TACO_LAUNCH_CHECK();

    struct Src { const float* W; int ld; int k0; };
    Src src[NSEG];
    src[SG_IN_S] = {w_sF, U, 0}; src[SG_IN_CP] = {w_cpF, U, 0};
    for (int i = 0; i < 3; ++i) {
        src[SG_G_X0 + i] = {w->gru_Wg[i], 2 * U, 0};  src[SG_G_H0 + i] = {w->gru_Wg[i], 2 * U, U};     // rows: x then h (TF 1.2 GRUCell)
        src[SG_C_X0 + i] = {w->gru_Wc[i], U, 0};      src[SG_C_RH0 + i] = {w->gru_Wc[i], U, U};
    }
    src[SG_Y] = {w->out_W, OUT, 0}; src[SG_QP] = {w_qp, 2 * AU, 0}; src[SG_PM] = {w->pre_W1, 256, 0}; src[SG_P2] = {w->pre_W2, 128, 0};
    for (int s = 0; s < NSEG; ++s) {
        const int64_t total = (int64_t)NS * sd[s].slice_floats;
        const int blocks = (int)((total + 255) / 256);
        pack_seg_kernel<<<blocks, 256, 0, stm>>>(src[s].W, src[s].ld, src[s].k0, sd[s].K, sd[s].FL, sd[s].nw, sd[s].kpw, s, NCY, OUT, packed + sd[s].g_off);
        TACO_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int taco_decoder_fwd(const taco_decoder_args* a, void* stream) {
    TACO_CHECK(a, "taco_decoder_fwd: NULL args");
    TACO_CHECK(a->weights != nullptr, "taco_decoder_fwd: weights (biases, attention_v) is NULL");
    const taco_decoder_weights& W = *a->weights;
    TACO_CHECK(a->B >= 1 && a->B <= BPAD, "taco_decoder_fwd: B=%d must be in [1,%d] per launch", a->B, BPAD);
    TACO_CHECK(a->T >= 1, "taco_decoder_fwd: T=%d", a->T);
    TACO_CHECK(a->Tx >= 1 && a->Tx <= 256, "taco_decoder_fwd: Tx=%d must be in [1, 256]", a->Tx);
    TACO_CHECK(a->r >= 1 && MF * a->r <= 512, "taco_decoder_fwd: r=%d unsupported (80r must be <= 512)", a->r);
    TACO_CHECK(a->packed && a->keys && a->values && a->text_length && a->y && a->align && a->workspace, "taco_decoder_fwd: NULL pointer");
    TACO_CHECK((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->packed) & 15) == 0,
               "taco_decoder_fwd: workspace / packed must be 16-byte aligned");
    if (a->mode != TACO_DEC_INFER) TACO_CHECK(a->mel != nullptr, "taco_decoder_fwd: teacher/sched mode needs mel");
    if (a->mode == TACO_DEC_SCHED) TACO_CHECK(a->sample_mask != nullptr, "taco_decoder_fwd: sched mode needs sample_mask");
    TACO_CHECK((a->keep1 == nullptr) == (a->keep2 == nullptr), "taco_decoder_fwd: keep1/keep2 must both be set or both NULL");
    cudaStream_t st = (cudaStream_t)stream;

    static DecParams P;                                  // (kept off the stack; a decoder launch is not re-entrant)
    memset(&P, 0, sizeof(P));
    int64_t tot, ws_total;
    build_seg_table(P.seg, &tot);
    build_ws_layout(P.buf, &ws_total);
    P.trace_off = ws_total;
    P.a = *a;
    P.OUT = MF * a->r;
    P.NCY = (P.OUT + NS - 1) / NS;
    P.Tq = (a->Tx + 3) / 4;            // quarter width; the last quarter is ragged when Tx % 4 != 0
    TACO_CHECK(P.Tq <= 64 && P.NCY <= 16, "taco_decoder_fwd: Tx/4 = %d > 64 or 80r/32 > 16", P.Tq);
    P.in_b = W.in_b; P.out_b = W.out_b; P.pre_b1 = W.pre_b1; P.pre_b2 = W.pre_b2; P.att_v = W.att_v;
    for (int i = 0; i < 3; ++i) { P.gru_bg[i] = W.gru_bg[i]; P.gru_bc[i] = W.gru_bc[i]; }
    const FusedTail ft = fused_tail(a->r, tot);
    P.b_qF = a->packed + ft.b_qF; P.b_p1F = a->packed + ft.b_p1F; P.b_inS = a->packed + ft.b_inS;

    // ---- the per-step program: one record per stage ----
    int n = 0;
    auto stage = [&](int kind, int st_id, int pre_seg, int pre_src, int pre_tagd, int pre_skip0, int on_seg, int on_seg2, int on_src, int on_tagd, int on_s0) {
        P.items[n++] = Item{kind, st_id, pre_seg, pre_src, pre_tagd, pre_skip0, on_seg, on_seg2, on_src, on_tagd, on_s0};
    };
    stage(IK_F2CP, ST_IN, SG_IN_S, B_S, 0, 1, SG_IN_CP, -1, B_CP, 1, 2);         // s(t-1) | [ctx(t-1) | p2(t)] (no s, no context at t = 0)
    for (int i = 0; i < 3; ++i) {
        const int x = (i == 0) ? B_Z : B_H0 + (i - 1);
        stage(IK_F4, ST_G0 + i, SG_G_H0 + i, B_H0 + i, 0, 1, SG_G_X0 + i, -1, x, 1, 0);           // h_i(t-1) | x
        stage(IK_F2, ST_C0 + i, SG_C_X0 + i, x, 1, 0, SG_C_RH0 + i, -1, B_RH0 + i, 1, 0);         // x (already consumed by the gates) | r*h
    }
    if (a->mode != TACO_DEC_INFER) stage(IK_PM, ST_PM, -1, 0, 0, 0, -1, -1, 0, 0, 0);             // teacher frame t+1 -> p1m(t+1), off-chain
    stage(IK_F4X2, ST_OQP, -1, 0, 0, 0, SG_Y, SG_QP, B_S, 1, 0);
    stage(IK_AP2, ST_AP2, -1, 0, 0, 0, -1, -1, 0, 0, 0);
    P.n_items = n;

    // ---- residency: off-chain segments in TMEM (256 columns per warp), on-chain segments in shared memory; when the
    //      keys/values leave too little shared memory (large Tx) further segments move to TMEM ----
    int off = 16 + 6400 + 512 + 384 + 256 + 704 + 1024 + 64 + 256;
    off = (off + 31) / 32 * 32;
    P.smem_kv_off = off;
    off += P.Tq * KV_LD + P.Tq * ENC;
    off = (off + 31) / 32 * 32;
    P.smem_res_off = off;
    const int budget = (227 * 1024) / 4 - off;
    const int tm_first[] = {SG_IN_S, SG_G_H0, SG_G_H1, SG_G_H2, SG_C_X0, SG_C_X1, SG_C_X2, SG_PM};
    const int sm_order[] = {SG_IN_CP, SG_G_X0, SG_C_RH0, SG_G_X1, SG_C_RH1, SG_G_X2, SG_C_RH2, SG_QP, SG_Y, SG_P2};
    int tcol_all = 0, tcol_hi = 0;       // TMEM columns used in every warp's 128-column window / in the windows of warps 8-15 only
    auto to_tmem = [&](int s) {
        const int cols = P.seg[s].kpw * P.seg[s].FL;
        if (P.seg[s].nw == NWARP) {
            const int c = tcol_all > tcol_hi ? tcol_all : tcol_hi;
            if (c + cols > 128) return false;
            P.seg[s].tmem_col = c; tcol_all = tcol_hi = c + cols;
        } else {
            if (tcol_hi + cols > 128) return false;
            P.seg[s].tmem_col = tcol_hi; tcol_hi += cols;
        }
        P.seg[s].smem_off = -1;
        return true;
    };
    for (int s : tm_first) TACO_CHECK(to_tmem(s), "taco_decoder_fwd: TMEM plan overflow");
    int res = 0;
    for (int s : sm_order) {
        if (P.seg[s].slice_floats <= budget - res) { P.seg[s].smem_off = res; res += P.seg[s].slice_floats; }
        else TACO_CHECK(s != SG_IN_CP && to_tmem(s), "taco_decoder_fwd: weights do not fit in shared + tensor memory (Tx=%d)", a->Tx);
    }
    P.smem_total_floats = off + res;
    const size_t smem_bytes = (size_t)P.smem_total_floats * 4;

    static bool configured = false;
    if (!configured) {
        TACO_CUDA(cudaFuncSetAttribute(decoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TACO_CUDA(cudaFuncSetAttribute(decoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(NCTA); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    // all 32 clusters (128 CTAs) must be co-resident: the kernel is a dataflow machine without a grid barrier
    static int max_clusters = -1;
    if (max_clusters < 0) {
        cudaLaunchConfig_t q = cfg; q.dynamicSmemBytes = 227 * 1024;
        int nc = 0;
        TACO_CUDA(cudaOccupancyMaxActiveClusters(&nc, (void*)decoder_kernel<false>, &q));
        max_clusters = nc;
    }
    TACO_CHECK(max_clusters >= NCTA / 4, "taco_decoder_fwd: only %d clusters of 4 CTAs can be co-resident on this device (need %d)", max_clusters, NCTA / 4);
    TACO_CUDA(cudaMemsetAsync(a->workspace, 0, (size_t)ws_total * 8, st));
    static const bool trace_build = getenv("TACO_TRACE") != nullptr;     // per-stage clock trace (scripts/dec_time.py)
    if (trace_build && a->step_ns) TACO_CUDA(cudaLaunchKernelEx(&cfg, decoder_kernel<true>, P));
    else                           TACO_CUDA(cudaLaunchKernelEx(&cfg, decoder_kernel<false>, P));
    ++g_taco_launches;
    return 0;
}
