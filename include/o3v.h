/*
 * o3v.h -- C ABI of libo3v.so: the B200 (sm_100a) implementation of Open-o3-Video's
 * RL policy-objective hot path.
 *
 * Every entry point replaces a piece of the reference's Python/PyTorch path; the
 * reference file:line it replaces is cited on each declaration (paths relative to
 * /root/reference/src/r1-v/src/open_r1/).  The reference has no FFI of its own (it is
 * pure Python); the binding a maintainer adds is the ctypes stub in INTEGRATION.md,
 * which is what `open-o3-video_b200/_lib.py` implements.
 *
 * Conventions
 *  - POD arguments only: raw DEVICE pointers, sizes, scalars and a CUDA stream handle
 *    (`void*`, i.e. cudaStream_t / CUstream; NULL = legacy default stream).
 *  - The caller owns every buffer (inputs, outputs, workspace).  The library never
 *    allocates or frees device memory and keeps no pointer past the call.  Outputs are
 *    fully overwritten unless an `accumulate` flag says otherwise.
 *  - All work is enqueued on `stream`; no call synchronises the device.
 *  - Return 0 on success; negative = O3V_ERR_* argument/shape/alignment/arch violation;
 *    positive = cudaError_t.  Nothing throws or aborts across this boundary.
 *  - sm_100 only: on any other device every compute entry returns
 *    O3V_ERR_UNSUPPORTED_ARCH (there is no fallback path).
 *  - bf16 tensors are row-major, 16-byte aligned, leading dimension a multiple of 8.
 */
#ifndef O3V_H_
#define O3V_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O3V_VERSION 100 /* major*10000 + minor*100 + patch */

#define O3V_OK 0
#define O3V_ERR_INVALID_ARG (-1)      /* null pointer, non-positive size, bad flag */
#define O3V_ERR_ALIGNMENT (-2)        /* pointer / leading dimension alignment */
#define O3V_ERR_UNSUPPORTED_ARCH (-3) /* device is not sm_100 */
#define O3V_ERR_WORKSPACE (-4)        /* workspace too small */
#define O3V_ERR_DRIVER (-5)           /* cuTensorMapEncodeTiled / driver entry point failed */
#define O3V_ERR_SHAPE (-6)            /* unsupported shape (e.g. H % 64 != 0) */
#define O3V_ERR_UNSUPPORTED_MODE (-7) /* the entry point does not exist for the tile mode selected with o3v_set_tunable */

int o3v_version(void);
const char* o3v_strerror(int code);
/* 0 if the CURRENT device is sm_100, else O3V_ERR_UNSUPPORTED_ARCH. */
int o3v_check_device(void);
/* Diagnostic knobs for bench sweeps (defaults are the shipped configuration):
 *   "cta_pair"   1 = one CTA per 128-row tile, 2 = cta_group::2 pairs (256-row tiles), all GEMMs;
 *                "cta_pair_fwd" / "cta_pair_bwd" set it for K1 / K2 only (defaults 1 / 2)
 *   "fwd_groups" vocab splits per token block in K1 (0 = auto)
 *   "max_ctas"   cap on the persistent grid (0 = all SMs)
 *   "hint_fwd_a" / "hint_fwd_b" / "hint_fwd_store" / "hint_bwd_a" / "hint_bwd_b"  L2 eviction hint of the TMA
 *                loads of operand A / B and of the logits store (0 = none, 1 = evict first, 2 = evict last) */
int o3v_set_tunable(const char* name, int value);
/* Test helper: num_ctas CTAs holding smem_bytes of shared memory each spin for `clocks` SM clocks on `stream` (occupies
 * SMs beside a persistent GEMM launched on another stream). */
int o3v_debug_occupy_sms(int32_t num_ctas, int32_t smem_bytes, int64_t clocks, void* stream);

/* Diagnostic: the tcgen05 GEMM template on an arbitrary problem, D[M,N] (+)= A . B^T with
 * bf16 operands.  a_mn / b_mn = 0: operand stored [rows, K] (K contiguous, "K-major");
 * = 1: stored [K, rows] (rows contiguous, "MN-major").  out: bf16 or fp32 [M, ld_out]. */
int o3v_debug_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                   int32_t a_mn, int32_t b_mn, void* out, int64_t ld_out, int32_t out_fp32,
                   int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * K3a  first-EOS mask.   Replaces trainer/grpo_trainer.py:590-596.
 *   eos_idx[n] = first t with ids[n,t] == eos_id, else Tc      (int64, bit-exact)
 *   mask[n,t]  = (t <= eos_idx[n])                             (int32)
 * ---------------------------------------------------------------------------------- */
int o3v_eos_mask(const int64_t* completion_ids, int64_t N, int64_t Tc, int64_t eos_id,
                 int64_t* eos_idx, int32_t* completion_mask, void* stream);

/* ------------------------------------------------------------------------------------
 * K1  fused lm_head GEMM + online log-softmax statistics + target gather.
 * Replaces trainer/grpo_trainer.py:375 (`model(...).logits`, i.e. transformers'
 * lm_head = nn.Linear(H, V, bias=False)) and :380-383 (log_softmax + gather) for one
 * vocabulary slice [v_offset, v_offset + V) of the head.
 *
 *   hidden  [T, H] bf16 (each row already paired with its NEXT token id, i.e. the
 *           caller applied grpo_trainer.py:376-377's shift)
 *   weight  [V, H] bf16 (rows v_offset.. of lm_head.weight)
 *   targets [T] int64 GLOBAL vocab ids
 *   stats   [3, T] fp32 out: row max m, sum exp(z - m), target logit (0 if the target
 *           is outside this slice).  The [T, V] logits never reach HBM unless
 *   logits  != NULL: then bf16 logits are also stored to logits[t * ld_logits + v]
 *           (needed only by the chunked backward, see o3v_lmhead_dlogits).
 * ---------------------------------------------------------------------------------- */
size_t o3v_lmhead_fwd_workspace_bytes(int64_t T, int64_t V, int64_t H);
int o3v_lmhead_fwd(const void* hidden, const void* weight, const int64_t* targets,
                   int64_t T, int64_t V, int64_t H, int64_t v_offset,
                   float* stats, void* logits, int64_t ld_logits,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Merge P partial statistics (vocab slices of one GPU or of several GPUs after an
 * all-gather) into per-token log-probs: parts [P, 3, T] ->
 *   lse[t] = M + log(sum_p s_p * exp(m_p - M)),  logp[t] = sum_p z_p - lse[t].
 * Deterministic (fixed p order).  Finishes grpo_trainer.py:381-382. */
int o3v_lmhead_merge_stats(const float* parts, int64_t P, int64_t T,
                           float* logp, float* lse, void* stream);

/* Same merge fused with the vocab-parallel exchange: part_ptrs[p] (a HOST array of P <= 16 device
 * pointers) is rank p's [3, row_stride] triple in peer-mapped memory (NVLink P2P / symmetric
 * memory); every rank loads all P triples directly over NVLink instead of all-gathering them
 * first.  The caller orders the launch after a cross-rank barrier on `stream`. */
int o3v_lmhead_merge_stats_peers(const float* const* part_ptrs, int64_t P, int64_t row_stride, int64_t T,
                                 float* logp, float* lse, void* stream);

/* One-shot all-reduce (sum) of a bf16 buffer replicated in peer-mapped memory: bufs[p] (HOST array of
 * P <= 16 device pointers) is rank p's copy.  Rank `rank` owns the 1/P slice of 16-byte vectors:
 * it loads that slice from every peer over NVLink, adds in fp32 in rank order (deterministic,
 * identical on all ranks) and stores the bf16 result into every peer's buffer.  The kernel uses no
 * shared memory and few registers so that its `num_ctas` CTAs co-reside with the persistent GEMM
 * CTAs: it is meant to run on a side stream while K2b computes.  The caller brackets it with
 * cross-rank barriers (all partial sums written before, all results visible after).
 * n_elems % 8 == 0. */
int o3v_allreduce_bf16_peers(void* const* bufs, int64_t P, int64_t rank, int64_t n_elems, int32_t num_ctas,
                             void* stream);

/* Reduce-scatter (sum) by pull over the same replicated buffers: out[i] = sum_p bufs[p][elem_offset + i] for
 * i in [0, n_elems), fp32 in rank order, one bf16 rounding, written to the LOCAL `out` only.  Each token owner
 * calls it for ITS rows of dHidden (the data-parallel layout only needs the reduce-scatter half of
 * the all-reduce: half the NVLink bytes, no write fan-out).  Same co-residency properties and the same barrier
 * bracketing as o3v_allreduce_bf16_peers.  elem_offset % 8 == 0, n_elems % 8 == 0, out 16-byte aligned. */
int o3v_reduce_scatter_bf16_peers(void* const* bufs, int64_t P, int64_t elem_offset, int64_t n_elems, void* out,
                                  int32_t num_ctas, void* stream);

/* ------------------------------------------------------------------------------------
 * K2  chunked fused backward of K1 (replaces the autograd backward of
 * grpo_trainer.py:375-383: softmax-backward + two GEMMs).
 *
 * o3v_lmhead_dlogits: in place on the bf16 logits stored by o3v_lmhead_fwd,
 *   P[t,v] = g[t] * ( [v + v_offset == targets[t]] - exp(z[t,v] - lse[t]) )
 * where g[t] = dLoss/dlogp[t] (from o3v_gspo_fwd_bwd).
 * o3v_lmhead_bwd_dhidden:  dH[T,H]  = P[T,V] . W[V,H]          (bf16 or fp32 out)
 * o3v_lmhead_bwd_dweight:  dW[V,H] (+)= P^T[V,T] . hidden[T,H]  (fp32, optional accumulate)
 * ---------------------------------------------------------------------------------- */
int o3v_lmhead_dlogits(void* logits, int64_t T, int64_t V, int64_t ld_logits,
                       const float* lse, const float* grad_logp, const int64_t* targets,
                       int64_t v_offset, void* stream);
int o3v_lmhead_bwd_dhidden(const void* dlogits, int64_t ld_dlogits, const void* weight,
                           int64_t T, int64_t V, int64_t H,
                           void* d_hidden, int32_t out_is_fp32, void* stream);
int o3v_lmhead_bwd_dweight(const void* dlogits, int64_t ld_dlogits, const void* hidden,
                           int64_t T, int64_t V, int64_t H,
                           float* d_weight, int32_t accumulate, void* stream);

/* out[i] += slabs[0][i] + slabs[1][i] + ... + slabs[S-1][i]  (fp32, slab order: deterministic).  The host layer splits
 * the K (token) range of the LAST vocabulary tiles of K2b over the SMs that would otherwise idle in the final, nearly
 * empty wave of the persistent grid: S plain o3v_lmhead_bwd_dweight launches on S streams write their partial dW tiles
 * into S slabs, this kernel folds them into dW.  n_elems % 4 == 0, 16-byte aligned pointers. */
int o3v_add_slabs_f32(const float* slabs, int64_t S, int64_t n_elems, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Exp-store backward (default path of round 2): no elementwise pass over the [T, V] chunk at all.
 *
 * o3v_lmhead_fwd_exp is o3v_lmhead_fwd (identical statistics, bit for bit) except that what it stores is
 *     E[t, v] = exp(z[t, v] - row_ref[t])   (bf16; rows with row_keep[t] == 0 store zeros, row_keep may be NULL)
 * with a per-row reference row_ref[t] chosen BEFORE the sweep (any value within ~80 of the row maximum: the wrapper
 * takes the maximum over a strided sample of 256 vocabulary rows, computed with this same kernel, plus 16).  Then
 * softmax[t, v] = E[t, v] * exp(row_ref[t] - lse[t]) is a PER-ROW rescale of E, and by linearity
 *     dH[t, :]  = a_t * (E . W)[t, :] + g_t * W[tcol_t, :]            a_t = -g_t * exp(row_ref[t] - lse[t])
 *     dW[v, :]  = (E^T . (a * hidden))[v, :] + sum_{t: tcol_t = v} g_t * hidden[t, :]
 * so both backward GEMMs read E as stored: the rescale and the one-hot term live in the K2a epilogue
 * (o3v_lmhead_bwd_dhidden_exp, optionally storing at the token owners as o3v_lmhead_bwd_dhidden_scatter does when
 * slot_ptrs != NULL), in a [T, H] pre-scale of the hidden states and in a deterministic scatter of T rows
 * (o3v_lmhead_bwd_dweight_exp: three launches).  Storing exp(z - ref) in bf16 is also MORE accurate than storing z
 * (relative error 2^-9 instead of |z| * 2^-9).
 *   rows       [T] 16-byte records from o3v_lmhead_softmax_rows (g, a, target column in the slice); sort_key [T]
 *              int64 (optional out): the caller sorts the tokens by it (stable) and passes the permutation as
 *   order      [T] int64, which makes the one-hot scatter deterministic (sequential fp32 sum per dW row)
 *   scaled_hidden  workspace, T * H bf16, 16-byte aligned
 * ---------------------------------------------------------------------------------- */
int o3v_lmhead_fwd_exp(const void* hidden, const void* weight, const int64_t* targets,
                       int64_t T, int64_t V, int64_t H, int64_t v_offset,
                       float* stats, void* expz, int64_t ld_expz, const float* row_ref, const int32_t* row_keep,
                       void* workspace, size_t workspace_bytes, void* stream);
int o3v_lmhead_softmax_rows(const float* lse, const float* grad_logp, const int64_t* targets, const float* row_ref,
                            int64_t v_offset, int64_t V, int64_t T, void* rows, int64_t* sort_key, void* stream);
int o3v_lmhead_bwd_dhidden_exp(const void* expz, int64_t ld_expz, const void* rows, const void* weight,
                               int64_t T, int64_t V, int64_t H, void* d_hidden, int32_t out_is_fp32,
                               void* const* slot_ptrs, int64_t P, int64_t rank, int64_t rows_per_owner,
                               int64_t slot_rows, int64_t row0, void* stream);
int o3v_lmhead_bwd_dweight_exp(const void* expz, int64_t ld_expz, const void* rows, const int64_t* order,
                               const void* hidden, int64_t T, int64_t V, int64_t H, float* d_weight,
                               int32_t accumulate, void* scaled_hidden, void* stream);

/* K2a fused with the reduce-scatter of the vocab-parallel path (dHidden partial sums go back to the token
 * owners).  Rank `rank` of P holds a vocabulary slice and computes partial dH for ALL token rows; row r of this call
 * is global token row0 + r, owned by rank (row0 + r) / rows_per_owner.  The epilogue stores every bf16 output tile
 * straight into the OWNER's slot buffer over NVLink (peer-mapped memory), slot `rank`:
 *     slot_ptrs[owner] + ((rank * slot_rows + (row0 + r) - owner * rows_per_owner) * H + h) * 2 bytes
 * so the transfer overlaps the GEMM tile by tile and no collective kernel runs beside it.  slot_ptrs: HOST array of
 * the P ranks' slot buffers [P, slot_rows, H] bf16.  After a cross-rank barrier (all K2a done) each owner sums its P
 * slots in rank order with o3v_sum_slots_bf16 (fp32 accumulate, deterministic, local HBM only). */
int o3v_lmhead_bwd_dhidden_scatter(const void* dlogits, int64_t ld_dlogits, const void* weight,
                                   int64_t T, int64_t V, int64_t H, void* const* slot_ptrs, int64_t P, int64_t rank,
                                   int64_t rows_per_owner, int64_t slot_rows, int64_t row0, void* stream);
int o3v_sum_slots_bf16(const void* slots, int64_t P, int64_t slot_rows, int64_t rows, int64_t H, void* out,
                       int32_t num_ctas, void* stream);

/* The same two GEMMs with the softmax backward FUSED into their operand pipeline: `logits` is the bf16 logits chunk
 * exactly as o3v_lmhead_fwd stored it (it is NOT modified), and every A tile is rewritten to P in shared memory
 * between its TMA load and the tcgen05.mma that reads it (by eight transform warps per CTA).  Replaces
 * o3v_lmhead_dlogits + o3v_lmhead_bwd_*: no elementwise pass over the chunk (2 x 2 x T x V bytes of HBM traffic).
 * `rows` [T] 16-byte records (o3v_lmhead_softmax_bwd_rows: per-token g, log2|g| - lse*log2e and the target column
 * within this vocabulary slice), caller-owned, 16-byte aligned, T * 16 bytes.  Needs the default tile mode
 * (cta_pair_bwd = 2, bwd_wide = 1), else O3V_ERR_UNSUPPORTED_MODE.  Rows with grad_logp == 0 (masked-out tokens)
 * are never read nor exponentiated. */
int o3v_lmhead_softmax_bwd_rows(const float* lse, const float* grad_logp, const int64_t* targets, int64_t v_offset,
                                int64_t V, int64_t T, void* rows, void* stream);
int o3v_lmhead_bwd_dhidden_fused(const void* logits, int64_t ld_logits, const void* rows, const void* weight,
                                 int64_t T, int64_t V, int64_t H, void* d_hidden, int32_t out_is_fp32, void* stream);
int o3v_lmhead_bwd_dweight_fused(const void* logits, int64_t ld_logits, const void* rows, const void* hidden,
                                 int64_t T, int64_t V, int64_t H, float* d_weight, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * K3  KL + group advantages + GSPO (or token-level) ratio / clip / loss, forward and
 * backward in one launch.  Replaces grpo_trainer.py:635-636, 658, 675-681, 691-706 and
 * the metrics at :711, :737.
 *
 * One call handles the sequences [seq_offset, seq_offset + n_seq) of a step of N sequences
 * (n_seq = N, seq_offset = 0 for the whole step at once; the chunked fused fwd+bwd calls it
 * once per token chunk, in ascending seq_offset order starting at 0, on the same workspace).
 *   logp, ref_logp [n_seq, Tc] fp32; old_logp [n_seq, Tc] fp32 or NULL (= logp.detach(),
 *   the reference's behaviour at :691); mask [n_seq, Tc] int32: rows of THIS call;
 *   rewards_per_func [N, F] fp32: the whole step; groups are contiguous blocks of G rows
 *   (`view(-1, G)`, :675).
 * Outputs (any may be NULL except loss):
 *   loss [1], mean_kl [1]: written by the call that completes sequence N (means over N);
 *   advantages [N], reward_std [N], completion_len [N] int32: indexed globally;
 *   grad_logp [n_seq, Tc] = dLoss/dlogp, per_token_kl [n_seq, Tc]: rows of this call.
 * workspace: o3v_gspo_workspace_bytes(N) bytes, the same buffer for every call of a step.
 * ---------------------------------------------------------------------------------- */
size_t o3v_gspo_workspace_bytes(int64_t N);
int o3v_gspo_fwd_bwd(const float* logp, const float* old_logp, const float* ref_logp,
                     const int32_t* mask, const float* rewards_per_func,
                     int64_t N, int64_t Tc, int64_t F, int64_t G,
                     int64_t seq_offset, int64_t n_seq,
                     float beta, float eps_low, float eps_high, int32_t gspo,
                     float* loss, float* mean_kl, float* advantages, float* reward_std,
                     int32_t* completion_len, float* grad_logp, float* per_token_kl,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K4  grounded rewards on parsed rollouts (struct of arrays, fp64).
 * Replaces the numeric cores of reward_func.py: ans_tiou_reward :128-143,
 * ans_viou_reward :210-226, thk_temporal_segment_reward :417-421,
 * thk_temporal_point_reward :453-467, thk_spatial_reward :490-525 and :542-603,
 * calculate_iou :356-386, convert_coord_format(_gqa) :337-354, with the task gating of
 * :99-103, :196, :396, :439, :484-531.  Text extraction stays in Python.
 *
 * Rollout r belongs to prompt q = r / G; ground truth is stored once per prompt.
 * out [R, 5] fp64 = (ans_tiou, ans_viou, thk_temporal_segment, thk_temporal_point,
 * thk_spatial).
 * ---------------------------------------------------------------------------------- */
#define O3V_TASK_VISUAL_QA 0          /* "visual QA" */
#define O3V_TASK_TEMPORAL_QA 1        /* "temporal QA" */
#define O3V_TASK_TEMPORAL_QA_MCQ 2    /* "temporal QA (MCQ)" */
#define O3V_TASK_TS_FREEFORM 3        /* "temporal-spatial free-form QA" */
#define O3V_TASK_GENERAL_MCQ 4        /* "General video QA MCQ" */
#define O3V_TASK_GENERAL_FREEFORM 5   /* "General video QA Free-form" */

#define O3V_RF_HAS_THINK 1   /* <think>..</think> matched            (reward_func.py:392) */
#define O3V_RF_HAS_ANSWER 2  /* <answer>..</answer> matched          (reward_func.py:482) */
#define O3V_RF_ANS_SEG 4     /* answer holds "<t>s</t>s to <t>e</t>s" (reward_func.py:119) */
#define O3V_RF_ANS_BOX 8     /* answer holds a <box> that parsed to 4 numbers (:212, :361) */
#define O3V_GF_VBOX 1        /* GT answer holds a <box>              (reward_func.py:204) */

/* first double of the slot of a box >= 32 that is NOT a list of 4 numbers (a quiet-NaN payload no parsed number has) */
#define O3V_INVALID_BOX_BITS 0x7FF8B0B0DEADBEEFull

typedef struct o3v_rewards_soa {
  int64_t R;  /* rollouts */
  int64_t G;  /* rollouts per prompt; GT arrays have Q = R / G rows */
  int32_t P;  /* max think timestamps per rollout */
  int32_t C;  /* max claims per rollout */
  int32_t Bc; /* max boxes per claim */
  int32_t Tb; /* max think boxes per rollout, visual-QA branch */
  int32_t K;  /* max key frames per prompt */
  int32_t O;  /* max objects per key frame */
  int32_t Gb; /* max GT boxes per object */
  int32_t pad_;
  /* per rollout */
  const int32_t* flags;       /* [R] O3V_RF_* */
  const double* ans_seg;      /* [R, 2] */
  const double* ans_box;      /* [R, 4] */
  const int32_t* n_times;     /* [R] */
  const double* think_times;  /* [R, P] */
  const int32_t* n_claims;    /* [R] (= len(parsed_claims), the divisor at :603) */
  const double* claim_t;      /* [R, C] */
  const int32_t* claim_nbox;  /* [R, C] */
  const uint32_t* claim_valid;/* [R, C] bit b: box b < 32 is a list of 4 numbers (:361); boxes b >= 32 (degenerate
                                 repetition loops) are valid unless their slot starts with O3V_INVALID_BOX_BITS */
  const double* claim_box;    /* [R, C, Bc, 4] pixels */
  const int32_t* n_tboxes;    /* [R] */
  const uint32_t* tbox_valid; /* [R] same convention */
  const double* think_box;    /* [R, Tb, 4] pixels */
  /* per prompt */
  const int32_t* task;        /* [Q] O3V_TASK_* */
  const double* step_percent; /* [Q] kwargs['step_percent'][0] of the prompt's batch (reward_func.py:431) */
  const int32_t* gt_flags;    /* [Q] O3V_GF_* */
  const double* gt_seg;       /* [Q, 2] */
  const double* gt_vbox;      /* [Q, 4] */
  const double* image_size;   /* [Q, 2] (W, H) */
  const double* image_refine; /* [Q, 2] */
  const int32_t* n_kf;        /* [Q] */
  const double* kf_time;      /* [Q, K] in key_frames order */
  const int32_t* n_obj;       /* [Q, K] */
  const int32_t* n_gtbox;     /* [Q, K, O] */
  const double* gt_box;       /* [Q, K, O, Gb, 4] normalised */
} o3v_rewards_soa;

int o3v_grounded_rewards(const o3v_rewards_soa* soa, double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * K5  V-STAR scorer numerics (the offline counterpart of the reward numerics).
 * Replaces, per result item and answer chain, eval/test/eval_vstar.py:90-109
 * (calculate_temporal_iou), :112-146 (compute_iou, calculate_bbox_iou) and :148-178
 * (calculate_spatial_metrics: per-GT-frame max IoU over the predicted boxes, their mean in
 * numpy's pairwise summation order, AP at IoU >= 0.1/0.3/0.5/0.7/0.9).
 * out [I, 14] fp64 = chain 1 (tIoU, mIoU, AP x5), chain 2 (same).  F <= 64, Pb <= 32.
 * ---------------------------------------------------------------------------------- */
typedef struct o3v_vstar_soa {
  int64_t I;                /* result items */
  int32_t F;                /* max GT boxes (annotated frames) per item */
  int32_t Pb;               /* max predicted boxes per frame */
  const int32_t* t_valid;   /* [I, 2] chain's temporal answer is a list of 2 numbers (:97-104) */
  const double* gt_seg;     /* [I, 2] item['timestamps'] */
  const double* pred_seg;   /* [I, 2, 2] */
  const int32_t* sp_valid;  /* [I, 2] chain's spatial answer is non-empty (:150, :293) */
  const int32_t* n_frames;  /* [I] */
  const double* gt_box;     /* [I, F, 4] xmin, ymin, xmax, ymax */
  const int32_t* n_pb;      /* [I, 2, F] predicted boxes for the frame (0: frame id absent / empty) */
  const uint32_t* pb_valid; /* [I, 2, F] bit b: box b is a list of 4 numbers (:114) */
  const double* pb;         /* [I, 2, F, Pb, 4] */
} o3v_vstar_soa;

int o3v_vstar_scores(const o3v_vstar_soa* soa, double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * K6  completion text -> the per-rollout arrays of o3v_rewards_soa (SURVEY.md 8f rank 1:
 * the step before K4).  Replaces, bit for bit, the regex / json / float() extraction of
 * reward_func.py: the <think> / <answer> spans (:91-93, :394, :437, :481-482), the answer's
 * "<t>a</t>s to <t>b</t>s" (:119-126, temporal tasks), the answer's first <box> (:211-223,
 * visual QA), every "<t>x</t>s" of <think> (:405-412, :447-449), the <box>es of <think>
 * (:492-511, visual QA) and parse_temporal_spatial_reasoning_process (:308-335, other tasks).
 *
 *   text     UTF-8 bytes of the R completions back to back, 16-byte aligned; the allocation
 *            must be readable up to the next multiple of 16 bytes past offsets[R]
 *   offsets  [R + 1] int64 byte offsets (offsets[0] = 0 is not required)
 *   task     [Q = R / G] O3V_TASK_* of each prompt (kwargs['task'][0] of its batch)
 *   P, C, Bc, Tb  capacities (>= 1) of the output rows (as in o3v_rewards_soa).  If a rollout holds
 *            more candidates than a capacity, overflow[i] reports a count that fits (0 = everything
 *            fitted) and that rollout's rows / counts are incomplete: re-run with larger rows.
 *            overflow = (think times, claims, boxes per claim, think boxes).
 *   outputs  rows are written only up to the counts (no zero fill); slots past a count may hold
 *            scratch values.
 * workspace: o3v_parse_workspace_bytes(R, P, C, Tb) bytes, 8-byte aligned (per-rollout records passed between
 * the three launches: span ends, candidate counts; and the dense work lists of the conversion launch).
 * ---------------------------------------------------------------------------------- */
typedef struct o3v_parse_args {
  int64_t R;
  int64_t G;
  int32_t P, C, Bc, Tb;
  const uint8_t* text;
  const int64_t* offsets;
  const int32_t* task;
  int32_t* flags;        /* [R] */
  double* ans_seg;       /* [R, 2] */
  double* ans_box;       /* [R, 4] */
  int32_t* n_times;      /* [R] */
  double* think_times;   /* [R, P] */
  int32_t* n_claims;     /* [R] */
  double* claim_t;       /* [R, C] */
  int32_t* claim_nbox;   /* [R, C] */
  uint32_t* claim_valid; /* [R, C] */
  double* claim_box;     /* [R, C, Bc, 4] */
  int32_t* n_tboxes;     /* [R] */
  uint32_t* tbox_valid;  /* [R] */
  double* think_box;     /* [R, Tb, 4] */
  int32_t* overflow;     /* [4] */
} o3v_parse_args;

size_t o3v_parse_workspace_bytes(int64_t R, int32_t P, int32_t C, int32_t Tb);
int o3v_parse_completions(const o3v_parse_args* args, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* O3V_H_ */
