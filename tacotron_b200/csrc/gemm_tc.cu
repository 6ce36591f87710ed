// gemm_tc.cu -- tcgen05 / TMA implementation of taco_linear_fwd (TACO_IMPL_TC), sm_100a.
//
// One kernel serves every feed-forward contraction of the model (conv bank as ONE grouped
// implicit GEMM, conv projections, highway, GRU input products, attention memory layer,
// post dense):  Y[(b,t), n] = epi( sum_j sum_c X[b, t+tap0+j, c] * W[j,c,n] ).
//
//  * A operand: activations stay in their NWC layout in HBM; a 3-D TMA tensor map (C, T, B)
//    with a {32 ch, 128 rows, 1} box loads the tile for tap j at time offset t0+tap0+j.
//    TMA out-of-bounds zero fill IS the TF 'same' padding (rows < 0 or >= T, per utterance),
//    and also pads C up to a multiple of 32 (C=80 in the post-net).
//  * B operand: weights pre-packed K-major [N][taps*Cpad], TF32-rounded (taco_pack_weight).
//  * both land in shared memory in the 128B-swizzled K-major canonical UMMA layout; one
//    elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8), accumulators in TMEM.
//  * bank mode: CTA n-tile = filter k (BN == bank_cout), K loop runs over that filter's k taps
//    only -- no zero-padded taps are multiplied.
//  * warp roles: warp0 TMA producer, warp1 MMA issuer + TMEM owner, warps2-5 epilogue
//    (tcgen05.ld 32x32b -> smem transpose -> fused epilogue -> coalesced 128B row stores).
//  * 96 KB of pipeline smem per CTA -> 2 CTAs/SM, so one CTA's epilogue overlaps the other's
//    main loop without a persistent scheduler.
//  * X3 = true (TACO_IMPL_TC3, precision 'fp32x3'): error-compensated 3xTF32 -- fp32-grade products on
//    the tensor cores.  Weights arrive pre-split (hi = TF32 head, lo = remainder: taco_pack_weight_x3),
//    the activation tile is split by the four otherwise idle epilogue warps INTO TENSOR MEMORY (lane = tile
//    row, 32 columns hi + 32 columns lo per stage) and every k-step issues lo.hi + hi.lo + hi.hi into the same
//    TMEM accumulator with the A operand read from TMEM.  Why TMEM: with both operands in shared memory a
//    128x128x8 TF32 MMA reads 8 KB per 64 cycles = the SM's whole 128 B/clk of shared-memory bandwidth, and
//    the TMA writes + the in-smem split of the previous version (another 96 KB per k-block) made every k-block
//    take ~1500 cycles instead of the 768 the tensor pipe needs (profiles/r02_ncu_summary.md: tensor pipe 34-57 %
//    active).  A in TMEM halves the MMA's shared-memory reads and removes the split's writes; the stage shrinks
//    to 48 KB, so two CTAs share an SM again and one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include "epilogue.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // floats: one 128-byte swizzle row
constexpr int A_STAGE_BYTES = TC_BM * TC_BK * 4;   // 16 KB

struct TcArgs {
    int B, T;
    int cchunks;       // round_up(C,32)/32
    int Cpad;
    int taps, tap0;    // dense / plain conv
    int bank;          // 1: taps = n_tile+1, tap0 = -(taps-1)/2
    int N;             // valid output columns (highway: 2U)
    int tiles_per_seq; // M tiles per utterance
    int tile_stride;   // 128, or 127 when pooling
    int pool;
    int lo_row0;       // X3: first row of the lo half in the packed weight buffer (= rows of the hi half)
    int direct;        // 1: register-direct vectorised epilogue (all row strides multiples of 4 floats), 0: transposing epilogue
    EpiParams e;
};

template <int BN, int STAGES, bool X3 = false>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + (X3 ? 2 : 1) * B_STAGE_BYTES;     // X3: A (raw fp32) | B_hi | B_lo
    static constexpr int ACC_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;   // TMEM accumulator columns
    static constexpr int A_COLS = X3 ? 2 * TC_BK : 0;                 // X3: per stage 32 columns A_hi + 32 columns A_lo in TMEM
    static constexpr int NEED_COLS = ACC_COLS + STAGES * A_COLS;
    static constexpr int TMEM_COLS = NEED_COLS <= 32 ? 32 : NEED_COLS <= 64 ? 64 : NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
    static_assert(NEED_COLS <= 512, "tensor memory: accumulator + A stages exceed 512 columns");
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = PIPE_BYTES;
    static constexpr int CONST_OFFSET = PIPE_BYTES + 128;   // direct epilogue: bias[256] | scale[256] | shift[256] | first rows [4][32]
    static constexpr int TOTAL = PIPE_BYTES + 128 + 3072 + 512 + 1024;   // barriers + epilogue constants + alignment slack
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B canonical layout: 8-row groups 1024 B apart (SBO), LBO unused.
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 14 bits
    d |= (uint64_t)1 << 16;                           // leading byte offset (ignored for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

template <int BN>
__device__ __forceinline__ uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                 // accumulator F32
    d |= 2u << 7;                 // A = TF32
    d |= 2u << 10;                // B = TF32
    d |= (uint32_t)(BN >> 3) << 17;
    d |= (uint32_t)(TC_BM >> 4) << 24;
    return d;
}

// fp32 -> nearest TF32-representable value (ties away from zero), two integer instructions
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// MODE 0: normal epilogue (optional fused max-pool), MODE 1: highway (BN = 2U = 256)
template <int BN, int STAGES, int MODE, bool X3>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, TcArgs a) {
    using L = SmemLayout<BN, STAGES, X3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* conv_bar = tmem_full + 1;                  // X3: stage split into hi / lo by the converter warps
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(conv_bar + STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates
    const int n_tile = blockIdx.x;
    const int m_tile = blockIdx.y;
    const int b = m_tile / a.tiles_per_seq;
    const int t0 = (m_tile % a.tiles_per_seq) * a.tile_stride;
    const int n0 = n_tile * BN;
    int taps = a.taps, tap0 = a.tap0;
    if (a.bank) { taps = n_tile + 1; tap0 = -((taps - 1) / 2); }
    const int n_iters = taps * a.cchunks;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&conv_bar[s], 128); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        constexpr uint32_t cols = L::TMEM_COLS;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* As = smem + s * L::STAGE_BYTES;
                uint8_t* Bs = As + A_STAGE_BYTES;
                const int j = it / a.cchunks;
                const int c0 = (it - j * a.cchunks) * TC_BK;
                mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + (X3 ? 2 : 1) * L::B_STAGE_BYTES);
                tma_load_3d(As, &tmA, &full_bar[s], c0, t0 + tap0 + j, b);
                tma_load_2d(Bs, &tmB, &full_bar[s], j * a.Cpad + c0, n0);
                if (X3) tma_load_2d(Bs + L::B_STAGE_BYTES, &tmB, &full_bar[s], j * a.Cpad + c0, a.lo_row0 + n0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc<BN>();
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(X3 ? &conv_bar[s] : &full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
                const uint64_t adesc = make_smem_desc(a_addr);
                const uint64_t bdesc = make_smem_desc(a_addr + A_STAGE_BYTES);
                const uint64_t blo = make_smem_desc(a_addr + A_STAGE_BYTES + L::B_STAGE_BYTES);
                const uint32_t a_hi = tmem_base + (uint32_t)(L::ACC_COLS + s * L::A_COLS);      // X3: A operand in TMEM
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    // advance 8 tf32 along K: +32 bytes inside the swizzle atom (+2 in 16-byte units) / +8 TMEM columns
                    if (X3) {
                        tc_mma_tf32_ta(tmem_base, a_hi + TC_BK + 8 * k, bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);   // lo . hi
                        tc_mma_tf32_ta(tmem_base, a_hi + 8 * k, blo + (uint64_t)(2 * k), idesc, 1u);                                      // hi . lo
                        tc_mma_tf32_ta(tmem_base, a_hi + 8 * k, bdesc + (uint64_t)(2 * k), idesc, 1u);                                    // hi . hi
                    } else {
                        tc_mma_tf32(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                    (it > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit(&empty_bar[s]);      // frees the stage once these MMAs have read it
            }
            tc_commit(tmem_full);              // accumulator complete
        }
    } else {
        // ================= epilogue warps (2..5) =================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        if (X3) {
            // ---- during the main loop these 128 threads split each landed activation tile: x = hi + lo with hi = the
            //      nearest TF32 value (so that the result does not depend on how the tensor core narrows fp32) and lo = the
            //      exact remainder, itself rounded to TF32.  Lane = tile row = TMEM lane: the lane reads its 128-byte row out
            //      of the swizzled tile (8 lanes of a quarter-warp phase hit 8 different 16-byte chunks: conflict-free) and
            //      stores hi / lo into this stage's 64 TMEM columns ----
            const int row = q * 32 + lane;
            const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)L::ACC_COLS;
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                const uint8_t* arow = smem + s * L::STAGE_BYTES + row * 128;
                const uint32_t ta = lane_taddr + (uint32_t)(s * L::A_COLS);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int chunk = hf * 4 + c;
                        const float4 x = *reinterpret_cast<const float4*>(arow + ((chunk ^ (row & 7)) << 4));
                        // round to nearest (ties away) by integer arithmetic on the bit pattern: truncation would bias every
                        // term the same way and the bias grows linearly with K (1.6e-5 of max|ref| at K = 3072, measured)
                        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float hv = rn_tf32(xs[e]);
                            h[c * 4 + e] = __float_as_uint(hv);
                            l[c * 4 + e] = __float_as_uint(rn_tf32(xs[e] - hv));
                        }
                    }
                    tc_st_32x32b_x16(ta + (uint32_t)(hf * 16), h);
                    tc_st_32x32b_x16(ta + (uint32_t)(TC_BK + hf * 16), l);
                }
                tc_wait_st();
                tc_fence_before();                     // TMEM stores ordered before the arrive the MMA thread waits on
                mbar_arrive(&conv_bar[s]);
            }
        }
        if (a.direct) {
            // ================= register-direct epilogue =================
            // TMEM lane = output row: after tcgen05.ld a lane holds 32 consecutive columns of ITS row.  The fused epilogue
            // runs on those registers and the row segment leaves as eight 16-byte stores -- ~4 instructions per element
            // less than the transposing epilogue below (which measured ~37 K warp-instructions per 128x128 tile and made
            // every short-K contraction epilogue-bound: profiles/r02_ncu_summary.md).
            float* cst = reinterpret_cast<float*>(smem + L::CONST_OFFSET);
            float* frow = cst + 768;
            const int et = threadIdx.x - 64;
            for (int i = et; i < BN; i += 128) {
                if (MODE == 1) {
                    cst[i] = a.e.bias ? __ldg(a.e.bias + i) : 0.f;
                } else {
                    const int col = n0 + i;
                    const bool cv = col < a.N;
                    cst[i] = (cv && a.e.bias) ? __ldg(a.e.bias + col) : 0.f;
                    cst[256 + i] = (cv && a.e.scale) ? __ldg(a.e.scale + col) : 1.f;
                    cst[512 + i] = (cv && a.e.shift) ? __ldg(a.e.shift + col) : 0.f;
                }
            }
            named_bar_sync(1, 128);
            mbar_wait(tmem_full, 0);
            tc_fence_after();
            const int row = q * 32 + lane;
            const int t = t0 + row;
            const bool row_ok = (t < a.T) && (!a.pool || row < TC_BM - 1 || t == a.T - 1);
            const int64_t grow = (int64_t)b * a.T + t;
            constexpr int NCH = (MODE == 1) ? (BN / 2) / 32 : BN / 32;
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t v[32];
                float y[32];
                tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
                tc_wait_ld();
                const int c0 = ch * 32;                       // column offset inside the tile
                if (MODE == 1) {
                    uint32_t w[32];
                    tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN / 2 + ch * 32), w);
                    tc_wait_ld();
                    const float* hx = a.e.hx + grow * a.e.ldhx + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bh = *reinterpret_cast<const float4*>(cst + c0 + j);
                        const float4 bt = *reinterpret_cast<const float4*>(cst + BN / 2 + c0 + j);
                        const float4 x = row_ok ? __ldg(reinterpret_cast<const float4*>(hx + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float xs[4] = {x.x, x.y, x.z, x.w}, bhs[4] = {bh.x, bh.y, bh.z, bh.w}, bts[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float H = fmaxf(__uint_as_float(v[j + k]) + bhs[k], 0.f);
                            const float Tg = sigmoidf_acc(__uint_as_float(w[j + k]) + bts[k]);
                            y[j + k] = H * Tg + xs[k] * (1.0f - Tg);
                        }
                    }
                } else {
                    const int act = a.e.act;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 cb = *reinterpret_cast<const float4*>(cst + c0 + j);
                        const float4 sc = *reinterpret_cast<const float4*>(cst + 256 + c0 + j);
                        const float4 sh = *reinterpret_cast<const float4*>(cst + 512 + c0 + j);
                        const float cbs[4] = {cb.x, cb.y, cb.z, cb.w}, scs[4] = {sc.x, sc.y, sc.z, sc.w}, shs[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float u = __uint_as_float(v[j + k]) + cbs[k];
                            u = (act == TACO_ACT_RELU) ? fmaxf(u, 0.f) : (act == TACO_ACT_NONE) ? u : apply_act(u, act);
                            y[j + k] = fmaf(u, scs[k], shs[k]);
                        }
                    }
                    if (a.pool) {
                        // max_pooling1d(2,1,'same') over t: row t needs row t+1 = the next lane; lane 31 takes the first row of
                        // the next quarter from shared memory (tiles advance by 127 rows, so row 127 never needs a successor)
                        named_bar_sync(1, 128);               // previous chunk's first rows have been consumed
                        if (lane == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(frow + q * 32 + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                        }
                        named_bar_sync(1, 128);
                        const bool has_next = (t + 1 < a.T) && (row + 1 < TC_BM);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float nx = __shfl_down_sync(0xffffffffu, y[j], 1);
                            if (lane == 31) nx = frow[((q + 1) & 3) * 32 + j];
                            if (has_next) y[j] = fmaxf(y[j], nx);
                        }
                    }
                    if (a.e.keep) {
                        const uint4* kp = reinterpret_cast<const uint4*>(a.e.keep + grow * (int64_t)a.e.N + n0 + c0);
                        if (row_ok) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint4 kk = kp[h];
                                const uint32_t ks[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
                                for (int k = 0; k < 16; ++k)
                                    y[16 * h + k] = ((ks[k >> 2] >> (8 * (k & 3))) & 0xffu) ? y[16 * h + k] * a.e.keep_scale : 0.f;
                            }
                        }
                    }
                    if (a.e.residual && row_ok) {
                        const float* rp = a.e.residual + grow * a.e.ldr + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (n0 + c0 + j < a.N) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp + j));
                                y[j] += r4.x; y[j + 1] += r4.y; y[j + 2] += r4.z; y[j + 3] += r4.w;
                            }
                        }
                    }
                }
                if (row_ok) {
                    float* yp = a.e.Y + grow * a.e.ldy + (MODE == 1 ? 0 : n0) + c0;
                    const int ncols = (MODE == 1) ? BN / 2 : a.N;
                    const int cbase = (MODE == 1 ? 0 : n0) + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (cbase + j < ncols) *reinterpret_cast<float4*>(yp + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                }
            }
            tc_fence_before();
        } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        // scratch aliases the (now idle) pipeline buffers: [4 quarters][MODE?2:1][32][33] floats
        float* scratch_base = reinterpret_cast<float*>(smem);
        constexpr int SCR = 32 * 33;
        float* my_scr = scratch_base + q * (MODE == 1 ? 2 : 1) * SCR;
        const int64_t seq_row0 = (int64_t)b * a.T;
        constexpr int NCHUNK = (MODE == 1) ? (BN / 2) / 32 : BN / 32;
        constexpr int EPB = (MODE == 1) ? 8 : 16;     // epilogue rows per load batch (independent global loads in flight per lane)

        for (int ch = 0; ch < NCHUNK; ++ch) {
            uint32_t v[32];
            tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) my_scr[lane * 33 + j] = __uint_as_float(v[j]);
            if (MODE == 1) {
                tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN / 2 + ch * 32), v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) my_scr[SCR + lane * 33 + j] = __uint_as_float(v[j]);
            }
            if (a.pool) named_bar_sync(1, 128); else __syncwarp();

            const int col = n0 + ch * 32 + lane;     // MODE 1: n0 == 0, col in [0,U)
            int nrows = a.T - (t0 + q * 32);          // valid rows of this warp's quarter
            nrows = nrows < 0 ? 0 : (nrows > 32 ? 32 : nrows);
            const int64_t row0 = seq_row0 + t0 + q * 32;
            // The row loops below are written "all loads of a batch first, then math + stores" so that 8
            // independent global loads are in flight per lane (a naive row-at-a-time loop serialises on
            // L2 latency: measured 80 us of epilogue for a 2 us main loop).
            if (MODE == 1) {
                const int U = BN / 2;
                const float bH = a.e.bias ? __ldg(a.e.bias + col) : 0.f;
                const float bT = a.e.bias ? __ldg(a.e.bias + U + col) : 0.f;
                for (int rr0 = 0; rr0 < nrows; rr0 += EPB) {
                    float xin[EPB];
#pragma unroll
                    for (int i = 0; i < EPB; ++i)
                        xin[i] = (rr0 + i < nrows) ? __ldg(a.e.hx + (row0 + rr0 + i) * a.e.ldhx + col) : 0.f;
#pragma unroll
                    for (int i = 0; i < EPB; ++i) {
                        if (rr0 + i < nrows) {
                            const float H = fmaxf(my_scr[(rr0 + i) * 33 + lane] + bH, 0.f);
                            const float Tg = sigmoidf_acc(my_scr[SCR + (rr0 + i) * 33 + lane] + bT);
                            a.e.Y[(row0 + rr0 + i) * a.e.ldy + col] = H * Tg + xin[i] * (1.0f - Tg);
                        }
                    }
                }
            } else {
                const bool cvalid = col < a.N;
                const float cb = (cvalid && a.e.bias) ? __ldg(a.e.bias + col) : 0.f;
                const float csc = (cvalid && a.e.scale) ? __ldg(a.e.scale + col) : 1.f;
                const float csh = (cvalid && a.e.shift) ? __ldg(a.e.shift + col) : 0.f;
                const int act = a.e.act;
                auto affine = [&](float v) { return apply_act(v + cb, act) * csc + csh; };
                if (!a.pool && !a.e.keep && !a.e.residual) {
                    // fast path (the 1025-wide spectrogram projection lands here: its rows are not 16-byte aligned, so
                    // the register-direct epilogue cannot be used): pointer stepping, no per-element index arithmetic
                    if (cvalid) {
                        float* yp = a.e.Y + row0 * a.e.ldy + col;
                        const int64_t ys = a.e.ldy;
                        if (act == TACO_ACT_RELU) {
#pragma unroll 8
                            for (int rr = 0; rr < nrows; ++rr) { *yp = fmaf(fmaxf(my_scr[rr * 33 + lane] + cb, 0.f), csc, csh); yp += ys; }
                        } else if (act == TACO_ACT_NONE) {
#pragma unroll 8
                            for (int rr = 0; rr < nrows; ++rr) { *yp = fmaf(my_scr[rr * 33 + lane] + cb, csc, csh); yp += ys; }
                        } else {
#pragma unroll 4
                            for (int rr = 0; rr < nrows; ++rr) { *yp = affine(my_scr[rr * 33 + lane]); yp += ys; }
                        }
                    }
                } else if (!a.pool) {
                    if (cvalid) {
                        for (int rr0 = 0; rr0 < nrows; rr0 += EPB) {
                            float res[EPB];
                            float kp[EPB];
#pragma unroll
                            for (int i = 0; i < EPB; ++i) {
                                const bool ok = rr0 + i < nrows;
                                res[i] = (ok && a.e.residual) ? __ldg(a.e.residual + (row0 + rr0 + i) * a.e.ldr + col) : 0.f;
                                kp[i] = (ok && a.e.keep) ? (a.e.keep[(row0 + rr0 + i) * (int64_t)a.e.N + col] ? a.e.keep_scale : 0.f) : 1.f;
                            }
#pragma unroll
                            for (int i = 0; i < EPB; ++i) {
                                if (rr0 + i < nrows)
                                    a.e.Y[(row0 + rr0 + i) * a.e.ldy + col] = affine(my_scr[(rr0 + i) * 33 + lane]) * kp[i] + res[i];
                            }
                        }
                    }
                } else {
                    // fused max_pooling1d(2,1,'same') over t (models/ops.py:66-71): out[t] = max(e[t], e[t+1]),
                    // out[T-1] = e[T-1].  Tiles advance by 127 rows so row r+1 is always in this tile.
                    if (cvalid && nrows > 0) {
                        const float* nxt_scr = scratch_base + ((q + 1) & 3) * SCR;   // first row of the next quarter
                        float cur = affine(my_scr[lane]);
#pragma unroll 8
                        for (int rr = 0; rr < nrows; ++rr) {
                            const int r = q * 32 + rr;
                            const int t = t0 + r;
                            const bool has_next = (t + 1 < a.T) && (r + 1 < TC_BM);
                            float nxt = 0.f;
                            if (has_next) nxt = affine((rr < 31) ? my_scr[(rr + 1) * 33 + lane] : nxt_scr[lane]);
                            // rows 0..126 of the tile are produced here; row 127 only if it is the sequence end
                            if ((r < TC_BM - 1) || (t == a.T - 1))
                                a.e.Y[(seq_row0 + t) * a.e.ldy + col] = (t + 1 < a.T) ? fmaxf(cur, nxt) : cur;
                            cur = nxt;
                        }
                    }
                }
            }
            if (a.pool) named_bar_sync(1, 128); else __syncwarp();
        }
        tc_fence_before();
        }   // transposing epilogue
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        constexpr uint32_t cols = L::TMEM_COLS;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
    }
}

// ---- weight packing: TF [taps][C][N] -> K-major [N][taps*Cpad], TF32 round-to-nearest ----
__global__ void pack_weight_kernel(const float* __restrict__ W, int taps, int C, int N, int Cpad,
                                   float* __restrict__ dst, int64_t ld) {
    // one thread per (n, j, cpad-index); reads are strided by N but this runs once per weight update
    int64_t total = (int64_t)N * taps * Cpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cpad);
        int64_t r = i / Cpad;
        int j = (int)(r % taps);
        int n = (int)(r / taps);
        float v = 0.0f;
        if (c < C) {
            v = W[((int64_t)j * C + c) * N + n];
            uint32_t u;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
            v = __uint_as_float(u);
        }
        dst[(int64_t)n * ld + (int64_t)j * Cpad + c] = v;
    }
}

// X3: hi = TF32 round-to-nearest head, lo = (W - hi) cut to TF32 (the remainder has <= 13 significant bits; the cut drops
// at most 2^-22 |W|) -- both exactly representable, so the tensor core's own fp32 -> tf32 narrowing cannot change them
__global__ void pack_weight_x3_kernel(const float* __restrict__ W, int taps, int C, int N, int Cpad,
                                      float* __restrict__ dst_hi, float* __restrict__ dst_lo, int64_t ld) {
    int64_t total = (int64_t)N * taps * Cpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cpad);
        int64_t r = i / Cpad;
        int j = (int)(r % taps);
        int n = (int)(r / taps);
        float hi = 0.0f, lo = 0.0f;
        if (c < C) {
            const float v = W[((int64_t)j * C + c) * N + n];
            uint32_t u;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
            hi = __uint_as_float(u & 0xffffe000u);
            lo = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & 0xffffe000u);    // nearest, not cut: no systematic bias
        }
        dst_hi[(int64_t)n * ld + (int64_t)j * Cpad + c] = hi;
        dst_lo[(int64_t)n * ld + (int64_t)j * Cpad + c] = lo;
    }
}

// ---- driver entry point for cuTensorMapEncodeTiled (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}


// =====================================================================================================================
// Weight gradient of conv1d('same') / dense on the tensor cores (3xTF32, fp32-grade):
//     dW[j][c][n] += sum_{b,t} X[b, t+tap0+j, c] * dZ[b, t, n]
// The contraction runs over the ROWS of both operands (what a BLAS calls a TN product).  Neither operand is
// transposed in HBM and no MN-major descriptor is needed:
//   * a k-block = 32 consecutive rows t of one utterance.  TMA (3-D maps (C,T,B) / (N,T,B), no swizzle, OOB rows and
//     columns zero-filled = the 'same' padding and the ragged edges) lands X[32 rows][128 c] and dZ[32 rows][128 n];
//   * the four converter warps turn the X tile into the TMEM A operand: lane = c (the M index of the MMA), the 32
//     rows become 32 TMEM columns -- a conflict-free column read of the tile does the transpose for free -- split
//     into hi / lo TF32 halves (as in the forward kernel);
//   * the same warps transpose + split the dZ tile into two K-major SWIZZLE_128B tiles (hi, lo): thread = n reads 4
//     consecutive rows (conflict-free) and writes one 16-byte chunk per tile (8 lanes of a store phase hit 8
//     different chunks: conflict-free);
//   * one thread issues lo.hi + hi.lo + hi.hi tcgen05.mma (A from TMEM) per 8 rows; accumulator [128 c][128 n] in TMEM;
//   * split-K over the (b, t-block) list: grid.y CTAs each reduce a contiguous range and add their tile into dW with
//     vector reductions (red.global.add.v4.f32) -- same accumulate-with-atomics contract as taco_gemm's ta=1 path;
//   * the tensor core's fp32 accumulation is not round-to-nearest: over a 10 K-row chain the bias reached 8e-5 of
//     max|dW| (measured against float64).  Chains are therefore cut at 1024 rows: two TMEM accumulators alternate and
//     the converter warps add the finished one into dW while the MMAs fill the other (1024-row chains: <1e-5).
// =====================================================================================================================
constexpr int DW_STAGES = 3;
constexpr int DW_TILE_BYTES = 32 * 128 * 4;               // every staged tile is 16 KB
constexpr int DW_STAGE_BYTES = 4 * DW_TILE_BYTES;         // X raw | dZ raw | dZ hi (K-major) | dZ lo (K-major)
constexpr int DW_ACC_COLS = 256;                         // two accumulators [128 c][128 n], used alternately
constexpr int DW_DRAIN = 32;                             // k-blocks (1024 rows) summed in tensor memory before the partial tile is added to dW
constexpr int DW_SMEM = DW_STAGES * DW_STAGE_BYTES + 256 + 1024;

struct DwArgs {
    int T, C, N, taps, tap0;
    int c_tiles, n_tiles;
    int blocks_per_seq;      // ceil(T / 32)
    int total_blocks;        // B * blocks_per_seq
    int per_split;
    float* dW; int64_t ldw; int64_t tap_stride;
    int vec4;                // rows of dW 16-byte aligned
};

__global__ void __launch_bounds__(192) dw_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                    const __grid_constant__ CUtensorMap tmZ, DwArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + DW_STAGES * DW_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + DW_STAGES;
    uint64_t* conv_bar = empty_bar + DW_STAGES;
    uint64_t* acc_full = conv_bar + DW_STAGES;            // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles = a.c_tiles * a.n_tiles;
    const int j = blockIdx.x / tiles;
    const int rem = blockIdx.x - j * tiles;
    const int c0 = (rem / a.n_tiles) * 128;
    const int n0 = (rem % a.n_tiles) * 128;
    const int g0 = blockIdx.y * a.per_split;
    const int g1 = min(a.total_blocks, g0 + a.per_split);
    const int n_iters = g1 - g0;                                   // >= 1 (host)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmZ);
        for (int s = 0; s < DW_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&conv_bar[s], 128); }
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % DW_STAGES;
                const uint32_t ph = (it / DW_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * DW_STAGE_BYTES;
                const int g = g0 + it;
                const int b = g / a.blocks_per_seq;
                const int t = (g - b * a.blocks_per_seq) * 32;
                mbar_arrive_expect_tx(&full_bar[s], 2 * DW_TILE_BYTES);
                tma_load_3d(st, &tmX, &full_bar[s], c0, t + a.tap0 + j, b);
                tma_load_3d(st + DW_TILE_BYTES, &tmZ, &full_bar[s], n0, t, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc<128>();
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % DW_STAGES;
                const uint32_t ph = (it / DW_STAGES) & 1;
                mbar_wait(&conv_bar[s], ph);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * DW_STAGE_BYTES);
                const uint64_t bhi = make_smem_desc(st + 2 * DW_TILE_BYTES);
                const uint64_t blo = make_smem_desc(st + 3 * DW_TILE_BYTES);
                const uint32_t a_hi = tmem_base + (uint32_t)(DW_ACC_COLS + s * 64);
                const int chain = it / DW_DRAIN;
                const bool first = (it - chain * DW_DRAIN) == 0;
                const uint32_t acc = tmem_base + (uint32_t)((chain & 1) * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc_mma_tf32_ta(acc, a_hi + 32 + 8 * k, bhi + (uint64_t)(2 * k), idesc, (!first || k > 0) ? 1u : 0u);   // lo . hi
                    tc_mma_tf32_ta(acc, a_hi + 8 * k, blo + (uint64_t)(2 * k), idesc, 1u);                                  // hi . lo
                    tc_mma_tf32_ta(acc, a_hi + 8 * k, bhi + (uint64_t)(2 * k), idesc, 1u);                                  // hi . hi
                }
                tc_commit(&empty_bar[s]);
                if (it + 1 == n_iters || (it + 1) % DW_DRAIN == 0) tc_commit(&acc_full[chain & 1]);   // this chain's accumulator is complete
            }
        }
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;                               // tile row of the accumulator (c) and dZ column (n) this thread converts
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int c = c0 + m;
        float* drow = a.dW + (int64_t)j * a.tap_stride + (int64_t)c * a.ldw + n0;
        // add the finished accumulator of chain `ch` into dW (its MMAs were committed to acc_full[ch & 1])
        auto drain = [&](int chn) {
            mbar_wait(&acc_full[chn & 1], (uint32_t)((chn >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[32];
                tc_ld_32x32b_x32(lane_taddr + (uint32_t)((chn & 1) * 128 + ch * 32), v);
                tc_wait_ld();
                if (c < a.C) {
                    float* dp = drow + ch * 32;
                    const int nb = n0 + ch * 32;
                    if (a.vec4) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            if (nb + i < a.N)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + i), "r"(v[i]), "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3]) : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < a.N) atomicAdd(dp + i, __uint_as_float(v[i]));
                    }
                }
            }
            tc_fence_before();
        };
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % DW_STAGES;
            const uint32_t ph = (it / DW_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            const float* xs = reinterpret_cast<const float*>(smem + s * DW_STAGE_BYTES) + m;            // [32 rows][128]
            const float* zs = xs + DW_TILE_BYTES / 4;
            uint8_t* zh = smem + s * DW_STAGE_BYTES + 2 * DW_TILE_BYTES + m * 128;
            const uint32_t ta = lane_taddr + (uint32_t)(DW_ACC_COLS + s * 64);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t h[16], l[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float x = xs[(hf * 16 + k) * 128];
                    const float hv = rn_tf32(x);
                    h[k] = __float_as_uint(hv);
                    l[k] = __float_as_uint(rn_tf32(x - hv));
                }
                tc_st_32x32b_x16(ta + (uint32_t)(hf * 16), h);
                tc_st_32x32b_x16(ta + (uint32_t)(32 + hf * 16), l);
            }
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
                float4 h4, l4;
                const float z0 = zs[(4 * kc + 0) * 128], z1 = zs[(4 * kc + 1) * 128], z2 = zs[(4 * kc + 2) * 128], z3 = zs[(4 * kc + 3) * 128];
                h4.x = rn_tf32(z0); l4.x = rn_tf32(z0 - h4.x);
                h4.y = rn_tf32(z1); l4.y = rn_tf32(z1 - h4.y);
                h4.z = rn_tf32(z2); l4.z = rn_tf32(z2 - h4.z);
                h4.w = rn_tf32(z3); l4.w = rn_tf32(z3 - h4.w);
                const int off = (kc ^ (m & 7)) << 4;
                *reinterpret_cast<float4*>(zh + off) = h4;
                *reinterpret_cast<float4*>(zh + DW_TILE_BYTES + off) = l4;
            }
            fence_proxy_async();                   // generic-proxy tile writes -> visible to the tensor core's async-proxy reads
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&conv_bar[s]);
            // a chain's last block has just been handed to the MMA thread: the PREVIOUS chain finished long ago -- add it to dW
            // now, while the tensor pipe works through this one.  (The next chain reuses that accumulator; its first block is
            // converted only after this drain, so the order needs no extra barrier.)
            if ((it + 1) % DW_DRAIN == 0 && it + 1 < n_iters && it / DW_DRAIN >= 1) drain(it / DW_DRAIN - 1);
        }
        // ---- the last two chains ----
        const int n_chains = (n_iters + DW_DRAIN - 1) / DW_DRAIN;
        if (n_chains >= 2) drain(n_chains - 2);
        drain(n_chains - 1);
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int BN, int STAGES, int MODE, bool X3 = false>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, dim3 grid, cudaStream_t st) {
    using L = SmemLayout<BN, STAGES, X3>;
    static bool configured = false;
    if (!configured) {
        TACO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, MODE, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    gemm_tc_kernel<BN, STAGES, MODE, X3><<<grid, 192, L::TOTAL, st>>>(tmA, tmB, a);
    TACO_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int taco_pack_weight_impl(const float* W, int taps, int C, int N, float* dst, int64_t ld_dst, cudaStream_t st) {
    const int Cpad = (C + 31) / 32 * 32;
    TACO_CHECK(ld_dst >= (int64_t)taps * Cpad, "taco_pack_weight: ld_dst %lld < taps*Cpad %d", (long long)ld_dst, taps * Cpad);
    int64_t total = (int64_t)N * taps * Cpad;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    pack_weight_kernel<<<blocks, 256, 0, st>>>(W, taps, C, N, Cpad, dst, ld_dst);
    TACO_LAUNCH_CHECK();
    return 0;
}

int taco_pack_weight_x3_impl(const float* W, int taps, int C, int N, float* dst_hi, float* dst_lo, int64_t ld_dst, cudaStream_t st) {
    const int Cpad = (C + 31) / 32 * 32;
    TACO_CHECK(ld_dst >= (int64_t)taps * Cpad, "taco_pack_weight_x3: ld_dst %lld < taps*Cpad %d", (long long)ld_dst, taps * Cpad);
    int64_t total = (int64_t)N * taps * Cpad;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    pack_weight_x3_kernel<<<blocks, 256, 0, st>>>(W, taps, C, N, Cpad, dst_hi, dst_lo, ld_dst);
    TACO_LAUNCH_CHECK();
    return 0;
}

int taco_conv_dw_tc_impl(float* dW, int64_t ldw, int64_t tap_stride, const float* X, int64_t ldx, const float* dZ, int64_t lddz,
                         int B, int T, int C, int N, int taps, int tap0, cudaStream_t st) {
    TACO_CHECK(taco_aligned16(X) && taco_aligned16(dZ), "taco_conv_dw: X / dZ must be 16-byte aligned");
    TACO_CHECK((ldx % 4) == 0 && (lddz % 4) == 0, "taco_conv_dw: ldx and lddz must be multiples of 4 floats (TMA 16B strides)");
    TACO_CHECK(B >= 1 && T >= 1 && C >= 1 && N >= 1 && taps >= 1, "taco_conv_dw: bad shape");
    EncodeTiledFn enc = get_encode_fn();
    TACO_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    DwArgs a;
    a.T = T; a.C = C; a.N = N; a.taps = taps; a.tap0 = tap0;
    a.c_tiles = (C + 127) / 128; a.n_tiles = (N + 127) / 128;
    a.blocks_per_seq = (T + 31) / 32;
    a.total_blocks = B * a.blocks_per_seq;
    const int tiles = a.c_tiles * a.n_tiles * taps;
    int splits = 148 / tiles;
    if (splits < 1) splits = 1;
    if (splits > a.total_blocks) splits = a.total_blocks;
    a.per_split = (a.total_blocks + splits - 1) / splits;
    splits = (a.total_blocks + a.per_split - 1) / a.per_split;        // every split owns at least one block
    a.dW = dW; a.ldw = ldw; a.tap_stride = tap_stride;
    a.vec4 = (taco_aligned16(dW) && (ldw % 4) == 0 && (tap_stride % 4) == 0) ? 1 : 0;
    CUtensorMap tmX, tmZ;
    {
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)ldx * 4, (cuuint64_t)T * (cuuint64_t)ldx * 4};
        cuuint32_t box[3] = {128, 32, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TACO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(X) failed: %d (C=%d T=%d B=%d ldx=%lld)", (int)r, C, T, B, (long long)ldx);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)lddz * 4, (cuuint64_t)T * (cuuint64_t)lddz * 4};
        cuuint32_t box[3] = {128, 32, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmZ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(dZ), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TACO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dZ) failed: %d (N=%d T=%d B=%d lddz=%lld)", (int)r, N, T, B, (long long)lddz);
    }
    static bool configured = false;
    if (!configured) {
        TACO_CUDA(cudaFuncSetAttribute(dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM));
        configured = true;
    }
    dw_tc_kernel<<<dim3((unsigned)tiles, (unsigned)splits), 192, DW_SMEM, st>>>(tmX, tmZ, a);
    TACO_LAUNCH_CHECK();
    return 0;
}

int taco_linear_tc(const taco_linear_desc* d, cudaStream_t st) {
    TACO_CHECK(d->Wp != nullptr, "taco_linear_fwd(TC): Wp (packed weights) is NULL");
    TACO_CHECK(taco_aligned16(d->X) && taco_aligned16(d->Wp), "taco_linear_fwd(TC): X / Wp must be 16-byte aligned");
    TACO_CHECK((d->ldx % 4) == 0 && (d->ldwp % 4) == 0, "taco_linear_fwd(TC): ldx and ldwp must be multiples of 4 floats (TMA 16B strides)");
    EncodeTiledFn enc = get_encode_fn();
    TACO_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    const int64_t M = (int64_t)d->B * d->T;
    if (M == 0 || d->N == 0) return 0;

    const bool highway = d->epilogue == TACO_EPI_HIGHWAY;
    int BN = highway ? 256 : 128;
    if (highway) TACO_CHECK(d->N == 256 && d->taps == 1 && d->bank_K == 0 && d->hx, "highway (TC): needs N == 256 (U = 128), dense, hx");
    if (d->bank_K > 0) TACO_CHECK(d->bank_cout == 128 && d->N == d->bank_K * 128, "bank (TC): bank_cout must be 128");
    if (d->pool) TACO_CHECK(!highway && !d->residual && !d->keep, "pool cannot be combined with highway / residual / dropout");

    TcArgs a;
    a.B = d->B; a.T = d->T;
    a.Cpad = (d->C + 31) / 32 * 32;
    a.cchunks = a.Cpad / 32;
    a.taps = d->taps; a.tap0 = d->tap0; a.bank = d->bank_K > 0 ? 1 : 0;
    a.N = d->N;
    a.pool = d->pool ? 1 : 0;
    const bool x3 = d->impl == TACO_IMPL_TC3;
    // few output tiles (encoder-side contractions: 32 utterances x 128 characters = 32 m-tiles): 32-column n-tiles put
    // four times as many CTAs on the machine, each with a 6-deep ring (the long-K conv projection 2048x3 -> 128 ran on
    // 32 of 148 SMs with a 2-deep ring: 187 us)
    {
        const int64_t mt = (int64_t)((d->T + TC_BM - 1) / TC_BM) * d->B;
        if (x3 && !highway && d->bank_K == 0 && !d->pool && mt * ((d->N + 127) / 128) <= 37 && d->N >= 64) BN = 32;
    }
    // register-direct epilogue: needs 16-byte aligned row segments everywhere it touches
    {
        const int nout = highway ? d->N / 2 : d->N;
        bool ok = (d->ldy % 4) == 0 && (nout % 4) == 0 && taco_aligned16(d->Y);
        if (d->residual) ok = ok && (d->ldr % 4) == 0 && taco_aligned16(d->residual);
        if (d->hx) ok = ok && (d->ldhx % 4) == 0 && taco_aligned16(d->hx);
        if (d->keep) ok = ok && (nout % 16) == 0 && taco_aligned16(d->keep);
        a.direct = ok ? 1 : 0;
    }
    a.lo_row0 = d->N;                                    // packed buffer = [hi rows 0..N) | lo rows N..2N)
    a.tile_stride = a.pool ? (TC_BM - 1) : TC_BM;
    a.tiles_per_seq = a.pool ? ((d->T - 1 + a.tile_stride - 1) / a.tile_stride) : ((d->T + TC_BM - 1) / TC_BM);
    if (a.tiles_per_seq < 1) a.tiles_per_seq = 1;
    a.e.Y = d->Y; a.e.ldy = d->ldy; a.e.bias = d->bias; a.e.scale = d->scale; a.e.shift = d->shift;
    a.e.keep = d->keep; a.e.keep_scale = d->keep_scale; a.e.residual = d->residual; a.e.ldr = d->ldr;
    a.e.hx = d->hx; a.e.ldhx = d->ldhx; a.e.act = d->act; a.e.N = highway ? d->N / 2 : d->N;

    const int max_taps = d->bank_K > 0 ? d->bank_K : d->taps;
    TACO_CHECK(d->ldwp >= (int64_t)max_taps * a.Cpad, "taco_linear_fwd(TC): ldwp too small for taps*Cpad");

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[3] = {(cuuint64_t)d->C, (cuuint64_t)d->T, (cuuint64_t)d->B};
        cuuint64_t strides[2] = {(cuuint64_t)d->ldx * 4, (cuuint64_t)d->T * (cuuint64_t)d->ldx * 4};
        cuuint32_t box[3] = {TC_BK, TC_BM, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d->X), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TACO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (C=%d T=%d B=%d ldx=%lld)", (int)r, d->C, d->T, d->B, (long long)d->ldx);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->ldwp, (cuuint64_t)d->N * (x3 ? 2u : 1u)};
        cuuint64_t strides[1] = {(cuuint64_t)d->ldwp * 4};
        cuuint32_t box[2] = {TC_BK, (cuuint32_t)BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d->Wp), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TACO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d (N=%d ldwp=%lld)", (int)r, d->N, (long long)d->ldwp);
    }
    dim3 grid((d->N + BN - 1) / BN, (unsigned)(a.tiles_per_seq * d->B));
    if (x3) {
        if (highway) return launch_tc<256, 2, 1, true>(tmA, tmB, a, grid, st);
        if (BN == 32) return launch_tc<32, 6, 0, true>(tmA, tmB, a, grid, st);
        // 2 stages x 48 KB + 256 TMEM columns per CTA: two CTAs per SM (4 loads in flight per SM, the epilogue of one under the
        // main loop of the other).  A 4-deep ring with one CTA per SM measured 2 % slower on the C2 step.
        return launch_tc<128, 2, 0, true>(tmA, tmB, a, grid, st);
    }
    if (highway) return launch_tc<256, 2, 1>(tmA, tmB, a, grid, st);
    // short K loops (<= 8 k-iterations: dense 128/256-wide inputs) are epilogue/latency bound: a 2-stage ring (64 KB)
    // lets three CTAs share an SM instead of two
    if (!a.bank && a.taps * a.cchunks <= 8) return launch_tc<128, 2, 0>(tmA, tmB, a, grid, st);
    return launch_tc<128, 3, 0>(tmA, tmB, a, grid, st);
}
