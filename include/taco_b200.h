/*
 * taco_b200.h -- C-ABI of libtaco_b200.so: the B200 (sm_100a) hot path of the Tacotron
 * mel/linear-spectrogram model (reference: barronalex/Tacotron, models/tacotron.py, models/ops.py).
 *
 * The reference has no FFI boundary of its own (it is 100% Python over TensorFlow 1.2); the
 * boundary a maintainer binds is therefore the set of TF-1.2 library calls its hot path makes.
 * Each entry point below names the reference call site(s) it replaces.  The Python host
 * (tacotron_b200/models/{ops,tacotron}.py) reaches these through ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - every function returns 0 on success, non-zero on error; taco_last_error() gives the
 *     message for the calling thread.  Nothing throws across the boundary.
 *   - all tensors are caller-owned DEVICE pointers (fp32 unless noted), contiguous in their
 *     last dimension, 16-byte aligned.  The library allocates nothing on the hot path; the
 *     caller passes workspaces where needed.
 *   - all launches are asynchronous on the caller's stream (cudaStream_t passed as void*).
 *   - activations are NWC: x[B][T][C] (TF layout); dense weights are W[in][out]; conv weights
 *     W[k][Cin][Cout]; GRU weights gates[in+n][2n] (r then u), candidate[in+n][n]  (TF 1.2).
 */
#ifndef TACO_B200_H
#define TACO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TACO_VERSION 100

enum { TACO_ACT_NONE = 0, TACO_ACT_RELU = 1, TACO_ACT_SIGMOID = 2, TACO_ACT_TANH = 3 };
enum { TACO_IMPL_TC = 0,   /* tcgen05 tensor cores, TF32 multiplies, fp32 accumulate (TMA-fed) */
       TACO_IMPL_SIMT = 1, /* fp32 FFMA, exact fp32 products                                  */
       TACO_IMPL_TC3 = 2   /* tcgen05 tensor cores, error-compensated 3xTF32 (x = hi + lo; lo.hi + hi.lo + hi.hi,
                              fp32 accumulate): fp32-grade products (~1e-6 relative).  Wp then holds the
                              hi rows [0,N) followed by the lo rows [N,2N) (taco_pack_weight_x3)          */ };
enum { TACO_EPI_NORMAL = 0,
       TACO_EPI_HIGHWAY = 1 /* N = 2U: cols [0,U) = H pre-act, [U,2U) = T pre-act; Y[M][U] =
                               relu(H)*sig(T) + X*(1-sig(T)), X = hx (models/ops.py:32-45)   */ };
enum { TACO_DEC_INFER = 0, TACO_DEC_TEACHER = 1, TACO_DEC_SCHED = 2 };

const char* taco_last_error(void);
int         taco_version(void);
/* Number of SMs of the current device and whether the TMA/tcgen05 path is usable (cc 10.x). */
int         taco_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * taco_linear_fwd -- one fused contraction + epilogue.  Replaces, depending on the fields:
 *   tf.layers.dense            models/ops.py:30,32,39  models/tacotron.py:40,42,148
 *   tf.layers.conv1d('same')   models/ops.py:54-62 (bank: all K filters in ONE call), :80-85
 *   tf.layers.batch_normalization (inference affine, folded into scale/shift)  ops.py:64,87
 *   tf.layers.dropout          models/tacotron.py:41,43   (explicit keep mask)
 *   residual add               models/ops.py:92
 *   highway gate               models/ops.py:32-45 (TACO_EPI_HIGHWAY)
 * Y[b,t,n] = epi( sum_{j<taps} sum_{c<C} X[b, t+tap0+j, c] * W[j,c,n] ),  zero outside [0,T).
 * epi(v) = (act(v + bias[n]) * scale[n] + shift[n]) * keep[b,t,n]*keep_scale + residual[b,t,n]
 * Bank mode (bank_K > 0): N = bank_K*bank_cout; output columns [(k-1)*cout, k*cout) come from
 * filter k (k taps, tap0 = -(k-1)/2).  W then points at the K filters stored back to back
 * (W1 [1][C][cout], W2 [2][C][cout], ...), Wp at the packed form made by taco_pack_weight.
 * ------------------------------------------------------------------------------------------ */
typedef struct taco_linear_desc {
    const float*   X;        /* [B][T][ldx] activations, C valid channels                    */
    int64_t        ldx;      /* floats between consecutive (b,t) rows                        */
    int32_t        B, T, C;
    int32_t        taps;     /* 1 = dense                                                    */
    int32_t        tap0;     /* -(taps-1)/2 for TF 'same'                                    */
    int32_t        N;        /* output channels (total)                                      */
    int32_t        bank_K;   /* 0, or number of bank filters                                 */
    int32_t        bank_cout;
    const float*   W;        /* TF layout [taps*C][N] row-major (SIMT path)                  */
    const float*   Wp;       /* packed K-major [N][ldwp] TF32-rounded (tensor-core path)     */
    int64_t        ldwp;     /* floats per packed row = max_taps * round_up(C,32)            */
    float*         Y;        /* [B*T][ldy]                                                   */
    int64_t        ldy;
    const float*   bias;     /* [N] or NULL                                                  */
    int32_t        act;      /* TACO_ACT_*                                                   */
    const float*   scale;    /* [N] or NULL   (BN: gamma/sqrt(var+1e-3))                     */
    const float*   shift;    /* [N] or NULL   (BN: beta - mean*scale)                        */
    const uint8_t* keep;     /* [B*T][N] dropout keep mask or NULL                           */
    float          keep_scale;
    const float*   residual; /* [B*T][ldr] or NULL                                           */
    int64_t        ldr;
    int32_t        epilogue; /* TACO_EPI_*                                                   */
    const float*   hx;       /* highway carry input X [B*T][ldhx]                            */
    int64_t        ldhx;
    int32_t        pool;     /* 1: fuse max_pooling1d(2,1,'same') over t after the affine    */
                             /*    (models/ops.py:66-71); tensor-core path only              */
    int32_t        impl;     /* TACO_IMPL_*                                                  */
} taco_linear_desc;

int taco_linear_fwd(const taco_linear_desc* d, void* stream);

/* Pack one TF-layout weight W[taps][C][N] into the tensor-core operand layout: row n of dst
 * (dst + (n)*ld_dst) holds, for tap j and channel c, W[j][c][n] at column j*round_up(C,32)+c,
 * rounded to TF32 (round-to-nearest); padding channels are zero.                             */
int taco_pack_weight(const float* W, int taps, int C, int N, float* dst, int64_t ld_dst, void* stream);
/* 3xTF32 operand pair for TACO_IMPL_TC3: dst_hi rows = TF32 round-to-nearest heads, dst_lo rows = the remainders
 * (W - hi, cut to TF32); same layout as taco_pack_weight.  For a bank, call once per filter with the row offsets
 * of that filter in the hi half and in the lo half (lo half starts at row bank_K*bank_cout).                 */
int taco_pack_weight_x3(const float* W, int taps, int C, int N, float* dst_hi, float* dst_lo, int64_t ld_dst, void* stream);

/* tf.layers.max_pooling1d(pool 2, stride 1, 'same') over t  (models/ops.py:66-71)            */
int taco_maxpool_fwd(const float* X, float* Y, int B, int T, int C, void* stream);

/* Embedding-style row gather with optional dropout keep mask:
 * Y[i][:] = table[ids[i]][:] * keep[i][:]*keep_scale   (tf.nn.embedding_lookup, tacotron.py:114) */
int taco_gather_rows(const float* table, const int32_t* ids, int rows, int width, int vocab,
                     const uint8_t* keep, float keep_scale, float* Y, void* stream);

/* values[b][j][:] = memory[b][j][:] if j < length[b] else 0   (BahdanauAttention _prepare_memory,
 * tacotron.py:48-52)                                                                          */
int taco_mask_rows(const float* X, const int32_t* length, float* Y, int B, int T, int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * taco_bigru_fwd -- tf.nn.bidirectional_dynamic_rnn(GRUCell(128), GRUCell(128)) without
 * sequence_length (models/ops.py:118-128).  The input-side products are hoisted out of the
 * loop by the caller (one taco_linear_fwd):
 *   xp[b][t][0:256]   = x.Wg_fw[0:128,:] + bg_fw     xp[b][t][256:384] = x.Wc_fw[0:128,:] + bc_fw
 *   xp[b][t][384:640] = x.Wg_bw[0:128,:] + bg_bw     xp[b][t][640:768] = x.Wc_bw[0:128,:] + bc_bw
 * wh_fw / wh_bw point at the h-side rows: Wg[128:256][256] followed by Wc[128:256][128] are
 * passed separately.  out[b][t][0:128] = fw state, [128:256] = bw state.
 * One persistent CTA per (utterance, direction), recurrent weights register-resident.
 * ------------------------------------------------------------------------------------------ */
int taco_bigru_fwd(const float* xp, const float* Wg_h_fw, const float* Wc_h_fw,
                   const float* Wg_h_bw, const float* Wc_h_bw,
                   float* out, int B, int T, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoder: tacotron.py:46-105 create_decoder + :136-138 dynamic_decode, i.e. per step
 *   pre_net(last frame) ++ attention -> InputProjection -> 3x GRUCell(256) -> Residual ->
 *   OutputProjection(80r) -> BahdanauAttention(query = cell output) -> attention layer.
 * The whole T-step loop runs in ONE persistent kernel: 128 co-resident CTAs in clusters of 4, activations exchanged as
 * {value, step tag} words through L2, weights resident in shared + tensor memory (csrc/decoder.cu).  Tx in [1,256], any
 * width (the four attention quarters are ragged when Tx % 4 != 0); B <= 32 per launch; 80 r <= 512.
 * ------------------------------------------------------------------------------------------ */
typedef struct taco_decoder_weights {   /* all TF layout, device pointers */
    const float *pre_W1, *pre_b1, *pre_W2, *pre_b2;      /* [80][256],[256],[256][128],[128]   */
    const float *in_W, *in_b;                            /* [384][256],[256]                   */
    const float *gru_Wg[3], *gru_bg[3], *gru_Wc[3], *gru_bc[3]; /* [512][512],[512],[512][256],[256] */
    const float *out_W, *out_b;                          /* [256][80r],[80r]                   */
    const float *att_Wq, *att_v, *att_Wa;                /* [80r][256],[256],[80r+256][256]    */
} taco_decoder_weights;

/* bytes of the packed-weight buffer / of the workspace taco_decoder_fwd needs */
size_t taco_decoder_packed_bytes(int r);
size_t taco_decoder_workspace_bytes(int B, int Tx, int T, int r);
/* re-lay the weights per CTA column slice (call once per weight update) */
int taco_decoder_pack(const taco_decoder_weights* w, int r, float* packed, void* stream);

typedef struct taco_decoder_args {
    const taco_decoder_weights* weights; /* HOST pointer; biases and attention_v are read from here */
    const float*   packed;       /* from taco_decoder_pack                                    */
    const float*   keys;         /* [B][Tx][256]  values . W_mem                              */
    const float*   values;       /* [B][Tx][256]  length-masked memory                        */
    const int32_t* text_length;  /* [B]                                                       */
    const float*   mel;          /* [B][T][80r] teacher inputs (TEACHER / SCHED) or NULL      */
    const uint8_t* sample_mask;  /* [T][B] 1 = feed own output (SCHED) or NULL                */
    const uint8_t* keep1;        /* [T][B][256] pre-net dropout keep masks or NULL            */
    const uint8_t* keep2;        /* [T][B][128]                                               */
    float          keep_scale;   /* 1/(1-rate)                                                */
    int32_t        mode;         /* TACO_DEC_*                                                */
    int32_t        B, Tx, T, r;
    float*         y;            /* [B][T][80r]   seq2seq_output                              */
    float*         align;        /* [B][T][Tx]    alignment history                           */
    void*          workspace;    /* taco_decoder_workspace_bytes                              */
    uint64_t*      step_ns;      /* [T] device buffer: %globaltimer at the start of each step, or NULL */
    float*         h_save;       /* [3][T][B][256] the three GRU state sequences (training: input of
                                    taco_decoder_bwd's batched recomputation), or NULL          */
} taco_decoder_args;

int taco_decoder_fwd(const taco_decoder_args* a, void* stream);

/* tacotron.py:156-160: partial[0] = sum|a-b| over n elements (deterministic two-stage).       */
int taco_l1_loss_fwd(const float* a, const float* b, int64_t n, float* partial_ws, float* out, void* stream);
int taco_l1_partial_count(void);   /* floats partial_ws must hold */

/* number of kernels this library has launched so far in this process (for bench.py's gpu_launches) */
unsigned long long taco_launch_count(void);

/* ==========================================================================================
 * Training path.  Replaces what `tf.gradients` + `tf.clip_by_global_norm` + `tf.train.AdamOptimizer`
 * build for Tacotron.add_train_op (models/tacotron.py:167-185) over the forward graph
 * (models/tacotron.py:107-165, models/ops.py:27-132).  The host side (tacotron_b200/models/grad.py)
 * is the hand-written reverse of the forward; the entry points below are its arithmetic.
 * The semantics of each are pinned by the function of the same name in tests/mirror_kernels.py.
 * ========================================================================================== */

/* C_z[M][N] = beta*C_z + opA_z[M][K] . opB_z[K][N]  for z < batch  (fp32 FFMA; beta in {0,1}).
 *   ta=0: A stored [M][K]: opA[m][k] = A[m + sh(k)][k mod kper],  sh(k) = shift + z*bshift + (k / kper)*dshift
 *   ta=1: A stored [K][M]: opA[m][k] = A[k + sh][m],              sh    = shift + z*bshift
 *         a shifted stored row that leaves its block of `period` rows (0: one block) reads as zero.
 *   tb=0: B stored [K][N];  tb=1: B stored [N][kper] per K-segment j at offset j*b_tap_stride.
 * This one contraction covers every dense / conv1d('same') data gradient (row shift = tap offset, zero fill =
 * the 'same' padding per utterance), every weight gradient (ta=1, batch = taps, split-K with atomics), the
 * h(t-1) products of the recurrent layers (shift = one step) and the per-utterance attention context GEMMs.  */
typedef struct taco_gemm_desc {
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float*       C; int64_t ldc;
    int32_t M, N, K;
    int32_t ta, tb;
    float   beta;
    int32_t shift, period;
    int32_t taps, dshift, kper;
    int64_t b_tap_stride;
    int32_t batch;
    int64_t a_bstride, b_bstride, c_bstride;   /* in elements */
    int32_t bshift;
} taco_gemm_desc;
int taco_gemm(const taco_gemm_desc* d, void* stream);
/* which kernel taco_gemm launches: 0 = exact-product FFMA (default), 1 = 3xTF32 mma.sync tensor cores (fp32-grade,
 * ~1e-6 relative; opt-in until it has had a hardware run).  Returns the previous setting. */
int taco_set_gemm_impl(int impl);
/* Weight gradient of tf.layers.conv1d('same') / tf.layers.dense on the tcgen05 tensor cores (error-compensated 3xTF32,
 * fp32-grade; models/ops.py:54,80, tacotron.py:40,42,148 under tf.gradients, tacotron.py:170):
 *     dW[j*tap_stride + c*ldw + n] += sum_{b<B, t<T} X[(b*T + t + tap0 + j)*ldx + c] * dZ[(b*T + t)*lddz + n]
 * for j < taps, c < C, n < N; rows outside [0,T) of an utterance read as zero (the 'same' padding).  Accumulates into dW
 * (split-K with atomic adds, like taco_gemm's ta=1 path).  X, dZ 16-byte aligned, ldx and lddz multiples of 4 floats. */
int taco_conv_dw(float* dW, int64_t ldw, int64_t tap_stride, const float* X, int64_t ldx, const float* dZ, int64_t lddz,
                 int32_t B, int32_t T, int32_t C, int32_t N, int32_t taps, int32_t tap0, void* stream);

/* out[n] += sum_m A[m][n] * (Bm ? Bm[m][n] - (R ? R[m][n] : 0) : 1)     (bias / batch-norm gradients) */
int taco_colsum(float* out, const float* A, int64_t lda, const float* Bm, int64_t ldb, const float* R, int64_t ldr,
                int M, int N, void* stream);
/* in place C = act(C + bias) (bias may be NULL); out = X * H[row+shift] (zero outside the period block) */
int taco_bias_act(float* C, int64_t ldc, int M, int N, const float* bias, int act, void* stream);
int taco_mul_shift(float* out, int64_t ldo, const float* X, int64_t ldx, const float* H, int64_t ldh, int M, int N,
                   int shift, int period, void* stream);
/* backward of taco_linear_fwd's epilogue: dZ = dY*gain*scale[n]*mask; relu mask = Y > 0 (no scale) or
 * (Y - R - shift[n])*scale[n] > 0.  dZ may alias dY.                                                       */
int taco_epi_bwd(float* dZ, int64_t lddz, const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* R,
                 int64_t ldr, int M, int N, int relu, const float* scale, const float* shift, float gain, void* stream);
/* in place dropout with an explicit keep mask [M][N]: X = keep ? X*gain : 0   (tacotron.py:41,43) */
int taco_epi_fwd_keep(float* X, int64_t ldx, const uint8_t* keep, int M, int N, float gain, void* stream);
/* dbeta += S1; dgamma += (S2 - beta*S1)/gamma   (tf.layers.batch_normalization, inference form, ops.py:64,87) */
int taco_bn_param_grad(float* dgamma, float* dbeta, const float* S1, const float* S2, const float* gamma,
                       const float* beta, int N, void* stream);
/* tf.layers.max_pooling1d(2,1,'same') backward (ops.py:66-71); ties go to the first element */
int taco_maxpool_bwd(float* dX, const float* dP, const float* X, int B, int T, int C, void* stream);
/* highway gate (ops.py:32-45) from saved pre-activations P = [h_pre | t_pre] ([M][2U]) and its backward */
int taco_highway_fwd(float* Y, int64_t ldy, const float* P, int64_t ldp, const float* X, int64_t ldx, int M, int U,
                     void* stream);
int taco_highway_bwd(float* dP, int64_t lddp, float* dXd, int64_t lddx, const float* dY, int64_t lddy, const float* P,
                     int64_t ldp, const float* X, int64_t ldx, int M, int U, void* stream);
/* dA = beta*dA + sign(A - B)    (tacotron.py:158-160) */
int taco_l1_bwd(float* dA, const float* A, const float* B, int64_t n, float beta, void* stream);
/* same, A / B as [rows][cols] contiguous and dA with a row pitch of ldd >= cols floats (a 16-byte-aligned pitch for the
 * [M][1025] spectrogram gradient, so that the tensor-core data / weight gradient kernels can read it through TMA) */
int taco_l1_bwd_ld(float* dA, int64_t ldd, const float* A, const float* B, int64_t rows, int64_t cols, float beta, void* stream);
/* dTable[ids[m]][:] += dRows[m][:]   (embedding_lookup backward, tacotron.py:111-114) */
int taco_scatter_add_rows(float* dTable, const int32_t* ids, const float* dRows, int rows, int width, int vocab,
                          void* stream);
/* serial part of the bidirectional-GRU backward (ops.py:118-128): see csrc/gru_bwd.cu */
int taco_bigru_bwd(float* dxp, const float* dOut, const float* out, const float* ACT, const float* Wg_h_fw,
                   const float* Wc_h_fw, const float* Wg_h_bw, const float* Wc_h_bw, int B, int T, void* stream);
/* decoder step inputs of the Training / ScheduledOutputTraining helpers, time-major: Xin [T][B][80], sel [T][B] */
int taco_dec_inputs(float* Xin, uint8_t* sel, const float* mel, const float* y, const uint8_t* sample_mask, int B, int T,
                    int r, int sched, void* stream);

/* serial part of the attention-decoder backward (tacotron.py:46-105,136-138): see csrc/decoder_bwd.cu.
 * All tensors time-major ([T][B][n]) unless noted; B <= 32.                                               */
typedef struct taco_decoder_bwd_args {
    int32_t B, T, Tx, r;
    float   keep_scale;
    const float *W_a, *W_q, *W_out, *W_in, *W1, *W2, *v;      /* TF layouts (see taco_decoder_weights) */
    const float *Wg[3], *Wc[3];
    const float *dy_ext;                  /* [T][B][80r] gradient arriving at seq2seq_output            */
    const float *RU[3], *C[3], *H[3];     /* [T][B][512] gates, [T][B][256] candidates, states          */
    const float *align;                   /* [B][T][Tx]                                                 */
    const float *values, *keys;           /* [B][Tx][256]                                               */
    const float *PQ, *PN1, *PN2;          /* [T][B][256] y.W_q ; pre-net activations [T][B][256|128]    */
    const uint8_t* sel;                   /* [T][B] 1 = step input was the model's own previous output  */
    float *DATT, *DY, *DPQ;               /* out: pre-activation gradients per step                     */
    float *DSCORE;                        /* out: [B][T][Tx]                                            */
    float *DCTX, *DG[3], *DC[3], *DZ, *DPN2, *DPN1, *DX;
    float *workspace;                     /* taco_decoder_bwd_workspace_bytes                           */
} taco_decoder_bwd_args;
size_t taco_decoder_bwd_workspace_bytes(void);
int    taco_decoder_bwd(const taco_decoder_bwd_args* a, void* stream);
/* dkeys[b][j][d] = sum_t DSCORE[b][t][j] v[d] (1-e^2); dv[d] += sum DSCORE e; e = tanh(keys + PQ[t][b]) */
int taco_attn_bwd_post(float* dkeys, float* dv, const float* DSCORE, const float* keys, const float* PQ, const float* v,
                       int B, int T, int Tx, void* stream);

/* tf.clip_by_global_norm(cap_grads) + tf.train.AdamOptimizer (tacotron.py:170-184) on the flat
 * parameter / gradient buffers: out = sum x^2 (deterministic); g *= clip/max(sqrt(sumsq), clip);
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m/(sqrt(v)+eps), lr_t = lr sqrt(1-b2^t)/(1-b1^t).   */
/* ==========================================================================================
 * Spectrogram inversion: the fused steps of audio.griffinlim between the two
 * library FFTs of an iteration (reference audio.py:67-97; reshape_frames inverse audio.py:30-35; librosa
 * stft/istft conventions).  n = frames = 4r*(T/4), F = n_fft/2+1 bins, L = hop*(n-1) samples; complex tensors
 * are interleaved (re, im) fp32.  Semantics: tests/mirror_kernels.py gl_*.
 * ========================================================================================== */
/* mag[b,f,k] = exp(spec[b, t(f), c(f)*F+k]*scale + shift) (scale/shift may be NULL), full = mag*exp(2 pi i u):
 * reshape_frames(forward=False) + the driver's de-normalisation (test.py:64) + np.exp + the random phase (:81) */
int taco_gl_init(float* full_c64, float* mag, const float* spec, const float* phase_u, int B, int T, int n, int r, int F,
                 const float* scale, const float* shift, void* stream);
/* tail of librosa.istft: window, overlap-add, window-sum-square normalisation, centre trim: fr [B][n][n_fft] -> y [B][L] */
int taco_gl_ola(float* y, const float* fr, int B, int n, int hop, int n_fft, int win_length, void* stream);
/* head of librosa.stft: reflect padding, framing, hann window: y [B][L] -> frw [B][n][n_fft] */
int taco_gl_frame(float* frw, const float* y, int B, int n, int hop, int n_fft, int win_length, void* stream);
/* full = mag * rebuilt/|rebuilt|  (audio.py:84,87) over `count` complex elements */
int taco_gl_phase(float* full_c64, const float* mag, const float* rebuilt_c64, int64_t count, void* stream);
/* the two transforms of a Griffin-Lim iteration (librosa.istft / librosa.stft inside audio.griffinlim, audio.py:84-86),
 * n_fft = 2048: X [rows][1025] complex (interleaved re, im) <-> x [rows][2048] real.  rfft is unnormalised, irfft scales
 * by 1/2048 and ignores the imaginary parts of the DC and Nyquist bins (numpy / cuFFT conventions).  Shared-memory
 * radix-2 Stockham transform, one CTA per row: no FFT library on the path.                                          */
int taco_rfft2048(float* X_c64, const float* x, int64_t rows, void* stream);
int taco_irfft2048(float* x, const float* X_c64, int64_t rows, void* stream);

/* Input data format: spectrograms are stored float16 (preprocess.py:179-180) and normalised in
 * that dtype (data_input.py:56-64) before the float32 cast (:38-39).  Bit-exact device version of those statements:
 * out[i][c] = f32( f16( f32( f16( f32(x) - f32(mean[c]) ) ) / std[c] ) ),  x [rows][W] float16, mean float16, std float32. */
int taco_normalize_f16(float* out, const void* x_f16, const void* mean_f16, const float* std_f32, int64_t rows, int W,
                       void* stream);

int taco_sumsq(const float* x, int64_t n, float* partial_ws, float* out, void* stream);
int taco_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
                   float clip, const float* sumsq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TACO_B200_H */
