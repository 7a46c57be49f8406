/* embodied_b200 -- C ABI of the B200-native actor-learner hot path.
 *
 * The reference (danijar/embodied) is 100% Python and has no FFI: its plugin
 * boundary is the duck-typed Agent / Env / Stream protocols
 * (embodied/core/base.py:1-73) plus Driver / Replay.  This header is the
 * boundary we put UNDERNEATH those Python classes: each entry point names the
 * reference function whose per-element work it replaces.  A maintainer of the
 * reference would bind it with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; emb_last_error() gives a
 *    thread-local message.
 *  - no allocation / ownership crosses the ABI: every buffer is a caller-owned
 *    DEVICE pointer (or pinned host pointer where stated) with explicit sizes.
 *  - every launch takes a cudaStream_t (as void*) and is asynchronous.
 *  - integer / byte paths are bit-exact with the reference's numpy code.
 */
#ifndef EMBODIED_B200_H_
#define EMBODIED_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMB_ABI_VERSION 1
#define EMB_MAX_KEYS 32

/* per-key row operation */
enum emb_op {
  EMB_OP_COPY = 0,        /* dst[row] = src[row]                                            */
  EMB_OP_FIRST = 1,       /* COPY, then byte 0 of the first row of every window := 1        */
                          /*   Replay._annotate_batch  embodied/core/replay.py:283-286      */
  EMB_OP_LAST = 2,        /* COPY | is_first(next row of the same window)                   */
                          /*   Replay._annotate_batch  embodied/core/replay.py:287-291      */
  EMB_OP_FILL32 = 3,      /* dst[row] = fill (one int32 per row): the 'consec' key          */
                          /*   streams.Consec.__next__ embodied/core/streams.py:134         */
  EMB_OP_MASK = 4,        /* dst[row] = src[row] * (elem)(!aux[row])  typed multiply        */
                          /*   Driver._mask            embodied/core/driver.py:72-74,84-87  */
  EMB_OP_NORM_U8_F32 = 5, /* COPY u8 rows, and dst2[row] = float(src)/255 - 0.5 (fp32)       */
                          /*   Encoder.__call__        dreamerv3/rssm.py:230                */
  EMB_OP_NOT = 6          /* dst[row] = !src[row] (bool rows)                               */
};

/* element type for EMB_OP_MASK */
enum emb_dtype {
  EMB_U8 = 0, EMB_BOOL = 1, EMB_I32 = 2, EMB_I64 = 3, EMB_F32 = 4, EMB_F64 = 5,
  EMB_F16 = 6, EMB_BF16 = 7, EMB_I16 = 8, EMB_I8 = 9, EMB_U16 = 10,
  EMB_U32 = 11, EMB_U64 = 12
};

/* One key (= one named field of a transition) of a row-copy launch.  Rows of a
 * key are `row_bytes` long and `*_stride` bytes apart. */
typedef struct emb_key {
  const void* src;      /* device; row r of the source is src + srow(r)*src_stride  */
  void* dst;            /* device; row r of the destination is dst + drow(r)*dst_stride */
  void* dst2;           /* EMB_OP_NORM_U8_F32: float32 output, dense rows (else NULL) */
  const void* aux;      /* EMB_OP_LAST: the is_first SOURCE table (stride aux_stride, indexed
                           like src);  EMB_OP_MASK: bool is_last[nrows] (stride 1, indexed by r) */
  uint64_t src_stride;
  uint64_t dst_stride;
  uint64_t dst2_stride; /* bytes between rows of dst2 */
  uint64_t aux_stride;
  uint32_t row_bytes;
  uint32_t op;          /* enum emb_op */
  uint32_t dtype;       /* enum emb_dtype (EMB_OP_MASK only) */
  int32_t fill;         /* EMB_OP_FILL32 */
} emb_key_t;

const char* emb_last_error(void);
int emb_abi_version(void);
/* Number of kernels this library has launched in this process (gpu_launches). */
uint64_t emb_launch_count(void);
/* Account for launches this library recorded into a CUDA graph that the caller
 * has just replayed (a replay does not pass through the entry points). */
void emb_launch_count_add(uint64_t n);
/* SM count of the current device (grid sizing), <0 on error. */
int emb_device_sm_count(void);

/* CUDA-event stopwatch for per-kernel timing on the launching stream.  Inside a
 * stream capture the record becomes an event-record node of the graph and is
 * re-recorded by every replay (bench.py reads it after the step's sync). */
int emb_event_create(void** ev);
int emb_event_record(void* ev, void* stream);
int emb_event_elapsed_ms(void* a, void* b, float* ms);
int emb_event_destroy(void* ev);

/* Diagnostic: stream `bytes` of `src` out of HBM with a persistent grid of `ncta`
 * CTAs (contiguous slice per CTA).  mode 0: the scan kernels' weight path -- TMA
 * bulk copies into an `nstages` x `stage_bytes` shared-memory ring; mode 1: plain
 * 16-byte loads.  Timed by the caller (tools/probe_read.py) to place the read-only
 * roofline of the scan next to the measured copy peak. */
int emb_probe_read(const void* src, uint64_t bytes, int32_t ncta, int32_t mode,
                   int32_t nstages, int32_t stage_bytes, void* sink, void* stream);

/* The row engine.  For r in [0, nrows):
 *     srow(r) = src_rows ? src_rows[r] : r        drow(r) = dst_rows ? dst_rows[r] : r
 *   rows with srow(r) < 0 or drow(r) < 0 are skipped (evicted chunk,
 *   Replay.update embodied/core/replay.py:146-149).
 * `window` = rows per sampled sequence (for EMB_OP_FIRST / EMB_OP_LAST); 0 if unused.
 * src_rows / dst_rows are DEVICE int64 arrays (or NULL = identity). */
int emb_rows_copy(const emb_key_t* keys, int nkeys,
                  const int64_t* src_rows, const int64_t* dst_rows,
                  int64_t nrows, int32_t window, void* stream);

/* Named entry points = emb_rows_copy restricted to the ops each reference
 * function performs (argument meaning identical). */

/* Replay._assemble_batch + _annotate_batch + streams.Consec contiguous copy
 * (embodied/core/replay.py:256-292, embodied/core/streams.py:131-138):
 * gather B windows of `window` rows into dense (B, window, ...) outputs. */
int emb_replay_gather(const emb_key_t* keys, int nkeys, const int64_t* src_rows,
                      int64_t nrows, int32_t window, void* stream);

/* Replay.add / Chunk.append for N workers at once
 * (embodied/core/replay.py:77-99, embodied/core/chunk.py:41-50). */
int emb_replay_append_rows(const emb_key_t* keys, int nkeys,
                           const int64_t* dst_rows, int64_t nrows, void* stream);

/* Replay.update / _setseq / Chunk.update
 * (embodied/core/replay.py:130-149,216-235, embodied/core/chunk.py:54-58). */
int emb_replay_scatter_update(const emb_key_t* keys, int nkeys,
                              const int64_t* dst_rows, int64_t nrows, void* stream);

/* Driver._step obs side: np.stack + cast/normalise for the policy, fused with
 * the replay append of the observation keys
 * (embodied/core/driver.py:65, dreamerv3/rssm.py:230). */
int emb_driver_stage_obs(const emb_key_t* keys, int nkeys,
                         const int64_t* dst_rows, int64_t nrows, void* stream);

/* Driver._step action side: mask actions where is_last, emit reset, and scatter
 * action / policy-output rows into the replay
 * (embodied/core/driver.py:72-76,84-87). */
int emb_driver_scatter_mask_actions(const emb_key_t* keys, int nkeys,
                                    const int64_t* dst_rows, int64_t nrows,
                                    void* stream);

/* Chunk.save / Chunk.load (embodied/core/chunk.py:64-99, driven by Replay.save / load,
 * replay.py:295-388): rows [row0, row0 + nrows) of every key's table <-> dense caller buffers
 * [nrows][row_bytes], one strided copy per key on `stream` (a slab's rows are consecutive table
 * rows, so no row list and no kernel are needed).
 *   export: keys[i].src = table base (src_stride = its row pitch), keys[i].dst = the buffer;
 *   import: keys[i].src = the buffer, keys[i].dst = table base (dst_stride = its row pitch).
 * The buffer may be device memory or (pinned) host memory; only row_bytes, src, dst and the two
 * strides of a key are read (a stride of 0 means row_bytes). */
int emb_replay_export_chunk(const emb_key_t* keys, int nkeys, int64_t row0, int64_t nrows,
                            void* stream);
int emb_replay_import_chunk(const emb_key_t* keys, int nkeys, int64_t row0, int64_t nrows,
                            void* stream);


/* ------------------------------------------------------------------ learner:
 * the fused RSSM scan (dreamerv3/rssm.py:61-92 `_observe`, :135-159 `_core`).
 * One persistent cooperative kernel walks the T steps; see
 * embodied_b200/csrc/rssm_fwd.cu for the phase schedule.  All activations are
 * TIME-MAJOR with the batch padded to 16 rows: [T][16][...]; rows >= B are
 * computed and ignored.  engine 1 = bf16 tensor-core weights (packed by
 * emb_rssm_pack), engine 0 = fp32 FFMA (parity). */
typedef struct emb_rssm_fwd_args {
  int32_t B, T, D, H, S, C, G;   /* batch rows (<=16), steps, deter, hidden, stoch, classes, blocks */
  int32_t engine;                /* 0 fp32 FFMA (parity), 1 bf16 mma + TMA weight ring,
                                  * 2 bf16 mma with register-staged weight loads (first version) */
  int32_t ncta;                  /* grid size the weights were packed for (<= SM count) */
  int32_t tma_cfg;               /* set by the library (ring stages | tiles << 8) */
  float unimix, eps;             /* 0.01, 1e-4 (embodied/jax/outs.py:210-216, nets.py:364) */
  /* packed weights.  bf16: per CTA a contiguous block [K/16][per][32 lanes][2 u32] of mma B
   * fragments, per = ceil((N/8) / ncta) n8 tiles, CTA c owning tiles [c*per, (c+1)*per);
   * fp32: whole layer [N/8][K][8].  Block-diagonal layers are stored as one [K][N] matrix
   * whose column tile decides the group. */
  const void* w_ph1;     /* [D][2H]: obs0/kernel[:D] | dynin0/kernel            rssm.py:83-85,141 */
  const void* w_logit;   /* [H][S*C]: obslogit/kernel                           rssm.py:168-171 */
  const void* w_hid;     /* [G][D/G+3H][D/G]: dynhid0/kernel                    rssm.py:149 */
  const void* w_gru;     /* [G][D/G][3D/G], columns reordered (j/8, gate, j%8)  rssm.py:152-154 */
  const void* w_in1;     /* dynin1/kernel row-major [S*C][H], bf16 or fp32      rssm.py:143 */
  const float *b0, *b1, *b_hid, *b_gru, *b_logit;   /* biases (flat, reference layout) */
  const float *s0, *s1, *s_hid, *s_obs;             /* rms-norm scales */
  /* inputs */
  const float* deter0;   /* [16][D]   carry */
  const void* x2;        /* silu(rms(dynin2(action))) (hoisted): fp32 [T][16][H]; bf16 engine:
                          * per step [H/16][32 lanes][8 bf16] mma A fragments               */
  const float* pre_tok;  /* [T][16][H]   tokens @ obs0[D:] + bias   (hoisted)  */
  const float* keep;     /* [T+1][16]    1 - reset; keep[T] = 1                */
  const float* gumbel;   /* [T][16][S*C] injected sampling noise               */
  /* outputs */
  float* deter;          /* [T][16][D] */
  float* logit;          /* [T][16][S*C] */
  int32_t* index;        /* [T][16][S]   sampled class of every latent */
  /* saved for the backward pass; y0[0], y1[0] are INPUTS (step 0, hoisted).
   * Only rows < B of `index` are written. */
  float* y0;             /* [T+1][16][H] pre-norm dynin0 */
  float* y1;             /* [T+1][16][H] pre-norm dynin1 */
  float* yhid;           /* [T][16][D]   pre-norm dynhid0 */
  float* gates;          /* [T][4][16][D] reset, cand, update after their nonlinearity; cand before tanh/reset */
  float* yobs;           /* [T][16][H]   pre-norm obs0 */
  float* sumsq;          /* [T][16]      row sums of yhid^2; ZEROED by the caller */
  float* probs;          /* [T][16][S*C] softmax(logit) before unimix; rows < B written */
  float* rstd;           /* [T+1][3][16] rsqrt(mean(y^2)+eps) of y0[t], y1[t], yobs[t] */
  /* scratch */
  void* deterA;          /* bf16 engines scratch: (2*16*D + 2*16*H) bf16, A-fragment order */
  uint32_t* barrier;     /* 64 ZEROED u32: [0] grid barrier; bf16 engine: [32] counts published x1 operands */
  uint64_t* timing;      /* optional [T][16] globaltimer marks of CTA 0 (NULL = off) */
  /* engine 1 only: the action branch is hoisted out of the scan entirely.  w_hid then
   * holds rows [deter_g | x0 | x1] of dynhid0 ([G][D/G+2H][D/G]) and the caller passes
   * hid_pre[T][16][D] = x2 @ dynhid0[g][D/G+2H:] + bias (one (B*T)-row GEMM); x2 is unused. */
  const float* hid_pre;
  float* sumsq_obs;      /* engine 1: [T][16] row sums of yobs^2; ZEROED by the caller */
} emb_rssm_fwd_args;

int emb_rssm_observe_fwd(const emb_rssm_fwd_args* args, void* stream);

/* 1 if engine 1 (bf16, TMA weight ring) of both scan kernels fits this model on `ncta` CTAs,
 * else 0 (reason in emb_last_error()); callers then use engine 2. */
int emb_rssm_tma_fits(int32_t D, int32_t H, int32_t S, int32_t C, int32_t G, int32_t ncta);
/* The same question for engine 2; if neither fits the caller runs the scan step by step. */
int emb_rssm_legacy_fits(int32_t D, int32_t H, int32_t S, int32_t C, int32_t G, int32_t ncta);

/* Back-propagation through time of the same scan (embodied_b200/csrc/rssm_bwd.cu).
 * Consumes the activations emb_rssm_observe_fwd saved and the upstream
 * gradients of its three outputs; produces, per step, the upstream gradient of
 * every in-scan layer, from which the caller forms all parameter gradients
 * with (B*T)-row GEMMs.  Weights are the TRANSPOSED matrices in the same packed
 * layouts as the forward pass. */
typedef struct emb_rssm_bwd_args {
  int32_t B, T, D, H, S, C, G;
  int32_t engine, ncta;
  int32_t hoist_x2;      /* 1: wt_hid has no action rows ([D/G][G*(D/G+2H)]); the caller forms g_x2
                          * (engine 1 requires it; the library's private copy re-uses the slot
                          * for its ring configuration) */
  float unimix, eps;
  const void* wt_in1;    /* [H][S*C]        dynin1/kernel^T                              */
  const void* wt_logit;  /* [S*C][H]        obslogit/kernel^T                            */
  const void* wt_ph1;    /* [2H][D]         (obs0/kernel[:D] | dynin0/kernel)^T          */
  const void* wt_gru;    /* [3D/G][D]       column (g, j) = dyngru/kernel[g][j][:]       */
  const void* wt_hid;    /* [D/G][G*(D/G+3H)] column (g, n) = dynhid0/kernel[g][n][:]    */
  const float *s0, *s1, *s_hid, *s_obs;
  /* saved by the forward pass */
  const float *keep, *deter0, *deter, *y0, *y1, *yobs, *yhid, *gates, *sumsq, *probs, *rstd;
  /* upstream gradients of the forward outputs, [T][16][..], rows >= B zero */
  const float *G_deter, *G_logit, *G_stoch;
  /* per-step layer gradients (outputs) */
  float* g_xo;           /* [T][16][H]    wrt silu(rms(obs0))                    */
  float* g_logit;        /* [T][16][S*C]  wrt obslogit's output (total)          */
  float* g_gates;        /* [T][16][3D]   wrt dyngru's output, (g, gate, j) order */
  float* g_h;            /* [T][16][D]    wrt silu(rms(dynhid0))                 */
  float* g_x0;           /* [T+1][16][H]  wrt silu(rms(dynin0)); ZEROED          */
  float* g_x1;           /* [T+1][16][H]  wrt silu(rms(dynin1)); ZEROED          */
  float* g_x2;           /* [T][16][H]    wrt the hoisted action branch; ZEROED  */
  /* scratch */
  float* g_stoch;        /* [16][S*C] */
  float* gd_carry;       /* [16][D] ZEROED; ends as the gradient wrt keep_0 * deter0 */
  float* gd_tmp;         /* [16][D] */
  float* dots;           /* [T+1][4][16] ZEROED: row dots of the norm backward of x0, x1, xo, h */
  uint32_t* barrier;     /* one ZEROED u32 */
  void* frag_scratch;    /* engine 1: 16 * (5*D + 4*H) bf16 -- operand fragments handed from
                          * one phase's epilogue to the next phase's TMA fetch */
  float* gx_part;        /* engine 1: [G][16][2H] fp32, ZEROED -- per-group partials of the
                          * gradients wrt x0 | x1 (summed by row CTAs, no atomics) */
  uint64_t* timing;      /* optional [T][16] globaltimer marks of CTA 0 (NULL = off) */
} emb_rssm_bwd_args;

int emb_rssm_observe_bwd(const emb_rssm_bwd_args* args, void* stream);

/* The KL / free-nats / entropy reduction of RSSM.loss (dreamerv3/rssm.py:120-133,
 * embodied/jax/outs.py:208-240) as one pass over the posterior and prior logits
 * each way (embodied_b200/csrc/kl.cu).  Logits: fp32 (dtype 0) or bf16 (1); row
 * (b, t) starts at element b*stride_b + t*stride_t, its S*C classes contiguous.
 * Outputs are dense fp32 [B*T] (fwd) / [B*T][S][C] (bwd).
 *   dyn = rep = max(KL(post || prior), free_nats) on the unimixed distributions,
 *   kl_raw the unclipped sum (the clip mask of the backward pass),
 *   ent_post / ent_prior the entropies summed over the S latents.
 * bwd: g_prior = g_dyn * dKL/dprior, g_post = g_rep * dKL/dpost (the two stop-
 * gradients of rssm.py:125-128), zero where kl_raw < free_nats. */
typedef struct emb_rssm_kl_args {
  const void* post;
  const void* prior;
  int32_t dtype_post, dtype_prior;
  int32_t B, T, S, C;
  int64_t post_stride_b, post_stride_t, prior_stride_b, prior_stride_t;
  float unimix, free_nats;
} emb_rssm_kl_args;

int emb_rssm_kl_fwd(const emb_rssm_kl_args* args, float* dyn, float* rep, float* kl_raw,
                    float* ent_post, float* ent_prior, void* stream);
int emb_rssm_kl_bwd(const emb_rssm_kl_args* args, const float* kl_raw, const float* g_dyn,
                    const float* g_rep, float* g_post, float* g_prior, void* stream);

/* Forward value of the straight-through one-hot sample (embodied/jax/outs.py:210-216,
 * 252-270) for the no-gradient paths: out[row][s*C + c] = (c == argmax_c(log(unimix(
 * softmax(logit[row][s]))) + gumbel[row][s][c])).  logit fp32 / bf16 with `logit_stride`
 * elements between rows, gumbel fp32 with `gumbel_stride` (0 = dense S*C), out fp32 / bf16
 * with `out_stride` (the imagination roll-out samples straight into its feature buffer);
 * index (optional) int32 [rows][S]. */
int emb_onehot_sample(const void* logit, int32_t dtype, int64_t logit_stride, const float* gumbel,
                      int64_t gumbel_stride, int64_t rows, int32_t S, int32_t C, float unimix,
                      void* out, int32_t out_dtype, int64_t out_stride, int32_t* index,
                      void* stream);

/* The lambda-return recurrence (dreamerv3/agent.py:482-490) as one launch, one thread
 * per row.  last / term / rew / boot: fp32 [rows][length]; ret: fp32 [rows][length-1]. */
int emb_lambda_return(const float* last, const float* term, const float* rew, const float* boot,
                      float* ret, int64_t rows, int32_t length, float disc, float lam, void* stream);

/* Gumbel(0, 1) noise for the categorical draws of one update (outs.OneHot / Categorical
 * sample = arg-max of log-probabilities + Gumbel noise; embodied/jax/outs.py:252-270 draws it
 * with jax.random.categorical): out fp32 [n], 16-byte aligned, one write-only pass.
 * Philox4x32-10 keyed by `seed` (callers pass a fresh seed per call); g = -log(E),
 * E = -log(U) clamped to [1e-7, 46]. */
int emb_gumbel_fill(float* out, int64_t n, uint64_t seed, void* stream);

/* The advantage recurrence of ppo_loss (ppo/agent.py:204-212), one launch, one thread per row:
 *   live = (1 - term)(1 - 1/hor), cont = (1 - last)(1 - term) lam,
 *   adv[t] = rew[t+1] + live[t+1] val[t+1] - val[t] + live[t+1] cont[t+1] adv[t+1],  tar = adv + val.
 * rew / val: fp32 [rows][length]; last / term: bool bytes [rows][length]; adv / tar: fp32
 * [rows][length-1].  Both outputs feed stop-gradient paths only. */
int emb_gae_advantage(const float* rew, const float* val, const uint8_t* last, const uint8_t* term,
                      float* adv, float* tar, int64_t rows, int32_t length, float hor, float lam,
                      void* stream);

/* ppo's optimiser chain on flat fp32 buffers (Agent._make_opt, ppo/agent.py:120-131; optax):
 * clip_by_global_norm(clip) -> scale_by_adam(b1, b2, eps) -> add_decayed_weights(wd, mask) ->
 * scale_by_learning_rate(linear_schedule(0, lr, warmup)).  Two passes over the gradient.
 * scratch: fp32 [scratch_len] block partials (>= 4 x SM count is enough); state: fp32 [2] on
 * the device = {updates applied so far (incremented here), global gradient norm of this call};
 * wdmask: fp32 0/1 per element, NULL when wd == 0. */
int emb_opt_clip_adam(float* master, const float* grad, float* mu, float* nu, const float* wdmask,
                      int64_t n, float* scratch, int32_t scratch_len, float* state, float lr,
                      int32_t warmup, float clip, float eps, float wd, float b1, float b2,
                      void* stream);

/* rms-norm (+ silu) over the last axis, one HBM pass each way
 * (embodied/jax/nets.py:361-399 Norm('rms') followed by act, eps 1e-4), with the
 * preceding layer's bias folded in:  y = act(rms_norm(x + bias) * scale).
 * x, y, gy, gx: [rows][cols] contiguous, dtype 0 = fp32 / 1 = bf16, 16-byte
 * aligned, cols % (16 / elem size) == 0; scale / gscale / bias / gbias fp32 [cols].
 * bias / gbias may be NULL (no bias); a bias needs cols <= 256 (the channel
 * axis of the convolutions -- dense layers add theirs in the GEMM epilogue).
 * bwd ADDS the scale / bias gradients into gscale / gbias and needs cols <= 2048. */
int emb_rmsnorm_act_fwd(const void* x, const float* scale, const float* bias, void* y,
                        int64_t rows, int32_t cols, int32_t dtype, int32_t act, float eps,
                        void* stream);
int emb_rmsnorm_act_bwd(const void* x, const float* scale, const float* bias, const void* gy,
                        void* gx, float* gscale, float* gbias, int64_t rows, int32_t cols,
                        int32_t dtype, int32_t act, float eps, void* stream);

/* The optimiser chain of dreamerv3 on the flat parameter buffer, two HBM passes
 * (dreamerv3/agent.py:342-379; embodied/jax/opt.py:109-164: clip_by_agc ->
 * scale_by_rms -> scale_by_momentum -> learning rate).  `chunks` cuts the flat
 * buffers into <= 4096-element pieces that lie inside one tensor (16-byte
 * aligned begins); `norms` is scratch [ntensors][2] = (|g|^2, |w|^2) and is
 * left filled (grad-norm metric).  hyper (DEVICE) = {lr, 1/(1-b1^t), 1/(1-b2^t),
 * b1, b2, eps, agc clip, pmin}. */
typedef struct emb_opt_chunk {
  int64_t begin;
  int32_t count;
  int32_t tensor;
} emb_opt_chunk;

int emb_opt_agc_rms_momentum(const float* grad, float* param, float* nu, float* mu,
                             const emb_opt_chunk* chunks, int32_t nchunks,
                             float* norms, int32_t ntensors, const float* hyper,
                             float* partials, const int32_t* tensor_first, void* stream);
/* Same update; additionally writes the new parameters rounded to bf16 into the flat
 * buffer `param_bf16` (same element offsets; NULL = skip) -- the compute-dtype copy
 * the next forward pass reads (embodied/jax/nets.py:243 casts parameters at use). */
int emb_opt_agc_rms_momentum_cast(const float* grad, float* param, float* nu, float* mu,
                                  void* param_bf16, const emb_opt_chunk* chunks, int32_t nchunks,
                                  float* norms, int32_t ntensors, const float* hyper,
                                  float* partials, const int32_t* tensor_first, void* stream);
/* partials: DEVICE scratch, 2 floats per chunk; tensor_first: DEVICE int32 [ntensors + 1], index of
 * every tensor's first chunk.  The per-tensor norms are summed chunk by chunk in a fixed order (no
 * atomics): data-parallel ranks that hold identical gradients compute bit-identical updates. */

/* One gradient bucket end to end on one stream (embodied/jax/opt.py:52-54 pmean + :109-164 chain):
 * ncclAllReduce(avg) of grad[elem_begin, elem_begin + elem_count) in place over `nccl_comm`
 * (an ncclComm_t created by the host; NULL = single process, no exchange), then
 * emb_opt_agc_rms_momentum_cast restricted to the bucket's tensors: `chunks` / `nchunks` are the
 * bucket's slice of the chunk table (chunk.begin / chunk.tensor stay GLOBAL offsets / indices),
 * norms[2*tensor_begin ... ) of its tensor_count tensors are reset and refilled.  Launched from
 * the backward pass as soon as the bucket's gradients are complete, on a side stream. */
int emb_allreduce_bucket_update(void* nccl_comm, float* grad, float* param, float* nu, float* mu,
                                void* param_bf16, int64_t elem_begin, int64_t elem_count,
                                const emb_opt_chunk* chunks, int32_t nchunks, float* norms,
                                int32_t tensor_begin, int32_t tensor_count, const float* hyper,
                                float* partials, const int32_t* tensor_first, int32_t chunk_begin,
                                void* stream);

/* Spatial glue of the dreamerv3 encoder / decoder on NHWC tensors, one HBM pass
 * each (dreamerv3/rssm.py:239-240 2x2 max-pool; :336,349 nearest x2 up-sampling).
 * (n, h, w, c) always describe the SMALL tensor (pool output / up-sample input);
 * dtype 0 = fp32 / 1 = bf16; c % (16 / elem size) == 0; 16-byte aligned.
 * idx: one u8 (fp32) / u16 (bf16) per 16-byte output vector, 2-bit argmax per
 * element (first maximum in (dy, dx) row-major order). */
int emb_maxpool2_nhwc_fwd(const void* x, void* y, void* idx, int64_t n, int32_t h, int32_t w,
                          int32_t c, int32_t dtype, void* stream);
int emb_maxpool2_nhwc_bwd(const void* gy, const void* idx, void* gx, int64_t n, int32_t h,
                          int32_t w, int32_t c, int32_t dtype, void* stream);
int emb_upsample2_nhwc_fwd(const void* x, void* y, int64_t n, int32_t h, int32_t w, int32_t c,
                           int32_t dtype, void* stream);
int emb_upsample2_nhwc_bwd(const void* gy, void* gx, int64_t n, int32_t h, int32_t w, int32_t c,
                           int32_t dtype, void* stream);

/* Forward-only fusions of the block-GRU core (dreamerv3/rssm.py:147-158) for the
 * no-gradient batched paths (imagination, policy); embodied_b200/csrc/gru.cu.  x / y /
 * pre are GROUPED [g][m][..] as the batched GEMMs produce them; scale / bias fp32 in
 * the reference's flat layout; dtype 0 = fp32 / 1 = bf16.
 *   rmsnorm_grouped: y[g][m][:] = act(rms_norm over the full row m of (x + bias) * scale)
 *   gru_gates: out[m][g*dg + j] = u * tanh(r * c) + (1 - u) * deter[m][g*dg + j] with
 *       (r, c, u) = pre[g][m][{0,1,2}*dg + j] + bias[g*3*dg + {0,1,2}*dg + j],
 *       r = sigmoid(r), u = sigmoid(u - 1).  deter / out rows are `deter_stride` /
 *       `out_stride` elements apart (0 = dense g*dg): the roll-out reads and writes the
 *       deter columns of its (rows, steps, deter | stoch) feature buffer in place. */
int emb_rmsnorm_grouped_fwd(const void* x, const float* scale, const float* bias, void* y,
                            int64_t m, int32_t g, int32_t dg, int32_t dtype, int32_t act, float eps,
                            void* stream);
int emb_gru_gates_fwd(const void* pre, const float* bias, const void* deter, void* out, int64_t m,
                      int32_t g, int32_t dg, int32_t dtype, int64_t deter_stride, int64_t out_stride,
                      void* stream);

/* The two thin 5x5 convolutions of dreamerv3 (3 image channels in: encoder layer 0,
 * dreamerv3/rssm.py:233-238; 3 channels out: decoder image head, rssm.py:349-352;
 * embodied/jax/nets.py:298-323 Conv2D SAME) as a skinny GEMM over all pixels plus one
 * HBM-bound rearrangement (embodied_b200/csrc/thinconv.cu).  NHWC, k odd, rows of
 * `out` / `z` have `kp` columns (multiple of 8, >= k*k*c); dtype 0 = fp32 / 1 = bf16.
 *   patches: out[p][(dy*k+dx)*c + ch] = x[p + sign*(dy-k/2, dx-k/2)][ch], zero outside
 *            the image and in the padding columns (sign = +1 forward, -1 = the
 *            backward of tapsum);
 *   tapsum:  y[p][ch] = bias[ch] + sum_{dy,dx} z[p + (dy-k/2, dx-k/2)][(dy*k+dx)*c + ch].
 * (n, h, w) always describe the OUTPUT grid.  up = 2 folds a nearest x2 up-sampling in
 * front of the convolution into the rearrangement: tapsum reads z on the (h/2, w/2)
 * grid at (p + offset) / 2; patches (sign = -1, its backward) reads x on the (2h, 2w)
 * grid and sums the 2x2 block of every output pixel. */
int emb_conv_patches_nhwc(const void* x, void* out, int64_t n, int32_t h, int32_t w, int32_t c,
                          int32_t k, int32_t kp, int32_t sign, int32_t up, int32_t dtype,
                          void* stream);
int emb_conv_tapsum_nhwc(const void* z, const float* bias, void* y, int64_t n, int32_t h, int32_t w,
                         int32_t c, int32_t k, int32_t kp, int32_t up, int32_t dtype, void* stream);

/* Weight packing for the RSSM scan kernels (embodied_b200/csrc/pack.cu): n8 column
 * tiles of fp32 matrices living in one flat parameter buffer (the in-scan layers of
 * dreamerv3/rssm.py:135-159, 81-86) -> per-CTA bf16 blocks in mma.m16n8k16 B-fragment
 * order, the layout rssm_fwd_tma.cu / rssm_bwd_tma.cu stream:
 *   dst[cta][ks_begin + kstep][tile][lane = nn*4 + kq][reg][half] =
 *       bf16(base[slot_off[s] + k*slot_kstride[s] + nn*slot_nstride[s]]),  s = cta*per + tile,
 *       k = kstep*16 + reg*8 + kq*2 + half, kstep in [0, ks_count); zeros where slot_off < 0.
 * dst holds ks_total k-steps per CTA; several calls with different (base offsets,
 * ks_begin) fill a matrix whose rows come from more than one tensor. */
int emb_pack_tiles(const float* base, const int64_t* slot_off, const int32_t* slot_kstride,
                   const int32_t* slot_nstride, void* dst, int64_t nslots, int32_t per,
                   int32_t ks_begin, int32_t ks_count, int32_t ks_total, void* stream);

/* Head losses (embodied_b200/csrc/losses.cu).
 * emb_twohot_loss_*: embodied/jax/outs.py:311-330 TwoHot.loss for fp32 logits [rows][nbins] and one
 * or two fp32 targets per row (target2 may be NULL; the critic's `ret` + slowreg * `slow value`
 * terms share one softmax, dreamerv3/agent.py:425-431,468-472):
 *   loss[r] = CE(twohot(target[r]), logits[r]) + weight2 * CE(twohot(target2[r]), logits[r]);
 *   lse[r]  = logsumexp(logits[r]) (saved for the backward pass);
 *   glogits = gloss[r] * ((1 + weight2) softmax - twohot(target) - weight2 twohot(target2)).
 * bins: fp32 [nbins], ascending (embodied/jax/heads.py:132-144); 2 <= nbins <= 1024.
 * emb_twohot_pred: TwoHot.pred (outs.py:285-309), the symmetric sum around the middle bin.
 * emb_loss_reduce: total = sum_i scales[i] * mean(terms[i]) and means[i] (dreamerv3/agent.py:237-240)
 * in one launch; terms / counts / scales are HOST arrays of n <= 16 entries (device pointers inside
 * `terms`); ticket: one zero-initialised device word the launch leaves at zero. */
int emb_twohot_loss_fwd(const float* logits, const float* target, const float* target2, float weight2,
                        const float* bins, float* loss, float* lse, int64_t rows, int32_t nbins,
                        void* stream);
int emb_twohot_loss_bwd(const float* logits, const float* target, const float* target2, float weight2,
                        const float* bins, const float* lse, const float* gloss, float* glogits,
                        int64_t rows, int32_t nbins, void* stream);
int emb_twohot_pred(const float* logits, const float* bins, float* pred, int64_t rows, int32_t nbins,
                    void* stream);
int emb_loss_reduce(const float* const* terms, const int64_t* counts, const float* scales, int32_t n,
                    float* means, float* total, uint32_t* ticket, void* stream);

/* The 5x5 SAME convolutions of the dreamerv3 encoder / decoder on 64..256 channels
 * (dreamerv3/rssm.py:233-240 conv -> pool -> norm -> act; :336-352 up-sample -> conv -> norm;
 * embodied/jax/nets.py:298-323 Conv2D, NHWC activations, HWIO kernels) as an implicit GEMM on
 * the tcgen05 tensor cores with TMEM accumulators and TMA (cp.async.bulk.tensor) operand tiles
 * (embodied_b200/csrc/conv_tc.cu).  bf16 in / out, fp32 accumulation.
 *   out[n][y][x][co] = bias[co] + sum_{ky,kx,ci} in[n][y+ky-k/2][x+kx-k/2][ci] * w_packed[ky*k+kx][co][ci]
 * with zeros outside the image.  w_packed: [k*k][cout][cin] bf16 (the HWIO kernel with its last
 * two axes swapped); packing the spatially flipped kernel with channels exchanged
 * ([24-tap][cin][cout]) makes the same launch the convolution's DATA GRADIENT.
 * Constraints: k in {1, 3, 5}; cin % 64 == 0; cout % 32 == 0, 32 <= cout <= 256; w <= 128;
 * 16-byte aligned pointers.  Any image size: an output tile is the largest run of whole rows (a
 * row count dividing h) or whole images (a count dividing n) of at most 128 pixels -- 128 where w
 * divides 128 and the rows pack, e.g. 96 pixels for 96-, 48- and 24-wide maps.  bias (fp32 [cout])
 * may be NULL. */
int emb_conv5x5_nhwc_tc(const void* in, const void* w_packed, const float* bias, void* out,
                        int64_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize,
                        void* stream);

/* General form of emb_conv5x5_nhwc_tc, used for the decoder's `nearest x2 up-sampling -> 5x5 conv`
 * stages (dreamerv3/rssm.py:336-352) in their SUB-PIXEL form: output pixel (2y+py, 2x+px) of such a
 * stage only sees a 3x3 neighbourhood of the low-resolution input, with the 5x5 taps that fall on
 * the same low-resolution pixel summed -- four 3x3 convolutions on the (h, w) grid, 36 instead of
 * 100 multiply-adds per low-resolution pixel and channel pair.  (h, w) is always the grid the GEMM's
 * pixels live on.
 *   out_up = 2: `out` is (n, 2h, 2w, cout); the launch writes phase out_phase = py*2+px of it.
 *   in_up  = 2: `in` is (n, 2h, 2w, cin); the reduction also runs over its four phases (phase-major
 *               weights w_packed[4*k*k][cout][cin]) -- the data gradient of the four phase
 *               convolutions in one launch. */
typedef struct emb_conv_tc_args {
  const void* in;
  const void* w_packed;
  const float* bias;
  void* out;
  int64_t n;
  int32_t h, w, cin, cout, ksize;
  int32_t in_up, out_up, out_phase;
} emb_conv_tc_args;
int emb_conv_nhwc_tc(const emb_conv_tc_args* args, void* stream);

/* Weight gradient of the same convolutions: dw[tap][m][n] += sum over pixels of
 * x[p + (ky-k/2, kx-k/2)][ci] * gy[p][co] (fp32, red.global.add: dw must hold zeros or the running
 * gradient).  The reduction axis is the pixel axis, so both NHWC tensors feed the tcgen05 MMAs
 * as MN-major operands straight from their TMA tiles.  m_is_in = 1: m = cin, n = cout (dw is the
 * HWIO gradient); 0: m = cout, n = cin (dw = its transpose per tap).  The M side must have 128 or
 * 256 channels; the N side any multiple of 8 <= 256 -- dw then has n_pad = 64 * ceil(n / 64) columns
 * (columns >= n stay zero: the out-of-bounds channels of its last slab are TMA zero fill).  Any image
 * size: the pixels are walked in chunks of whole rows / images (or pieces of a row) of <= 64 pixels.
 * ksize = 1 is the weight gradient of a plain matrix product over pixel rows. */
int emb_conv5x5_wgrad_tc(const void* x, const void* gy, float* dw, int64_t n, int32_t h, int32_t w,
                         int32_t cin, int32_t cout, int32_t ksize, int32_t m_is_in, void* stream);

/* General form: gy_up = 2 reads phase gy_phase = py*2+px of a gradient stored on the doubled grid
 * (n, 2h, 2w, cout) -- the weight gradient of one phase convolution of a sub-pixel decoder stage. */
typedef struct emb_conv_wgrad_tc_args {
  const void* x;
  const void* gy;
  float* dw;
  int64_t n;
  int32_t h, w, cin, cout, ksize;
  int32_t m_is_in, gy_up, gy_phase;
} emb_conv_wgrad_tc_args;
int emb_conv_wgrad_tc(const emb_conv_wgrad_tc_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* EMBODIED_B200_H_ */
