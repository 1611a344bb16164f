// K4: grounded (spatio-temporal) rewards on parsed rollouts, fp64, struct of arrays.
//
// Replaces the numeric cores of the reference's reward_func.py (see include/o3v.h for the
// line map).  HBM-bound: 16 lanes cooperate on one rollout (lane = prediction index, so a
// rollout's [P]/[C] rows are read with coalesced 128-byte requests); ground truth is shared
// by the G rollouts of a prompt and is served from L1/L2 after the first touch.
// Memory latency, not bandwidth, is what a rollout costs (a few hundred bytes behind a chain of
// dependent loads), so the loads are issued in TWO rounds: (1) every per-rollout / per-prompt scalar
// and the ground-truth staging copy, none of which depends on another load; (2) the lane's own
// timestamp, claim, claim boxes and think box, predicated on the counts and the task from round 1.
// The branches then run on registers and shared memory (round 1 used to be ~10 dependent round trips).
// Arithmetic follows the reference's float64 operation order; explicit __d*_rn intrinsics
// keep nvcc from contracting mul+add into FMA so that results are bit-identical to numpy's.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace o3v {

constexpr int kLanes = 16;             // lanes per rollout
// Threads per CTA (template parameter of the kernel; 24 resident warps per SM in every variant).  A warp holds two
// rollouts of one prompt, i.e. of one task, and only some tasks are expensive (the claim / proximity branches: ~7 us of
// dependent fp64 chains per warp against ~2 us), so a CTA keeps its SM slots until its slowest warp is done.  Measured
// at 65536 rollouts x 16 (tools/k4_ab.py, L2 flushed): 256 threads 70.5 us, 128 threads 67.5 us, 64 threads 81.9 us
// (the per-CTA ground-truth staging and launch cost outweigh the better packing).
constexpr int kRewardThreadsDefault = 128;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// reward_func.py:356-386 with boxA = GT, boxB = pred (validity of boxB checked by caller).
__device__ __forceinline__ double box_iou(const double a[4], const double b[4]) {
  const double xA = fmax(a[0], b[0]), yA = fmax(a[1], b[1]);
  const double xB = fmin(a[2], b[2]), yB = fmin(a[3], b[3]);
  const double inter = dmul(fmax(0.0, dsub(xB, xA)), fmax(0.0, dsub(yB, yA)));
  const double areaA = dmul(dsub(a[2], a[0]), dsub(a[3], a[1]));
  const double areaB = dmul(dsub(b[2], b[0]), dsub(b[3], b[1]));
  const double uni = dsub(dadd(areaA, areaB), inter);
  return uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
}

// boxes beyond the 32 bits of the validity mask carry their validity in the slot (include/o3v.h)
__device__ __forceinline__ bool box_slot_valid(const double* p) {
  return (unsigned long long)__double_as_longlong(*p) != O3V_INVALID_BOX_BITS;
}

__device__ __forceinline__ void load4(const double* p, double o[4]) {
  const double2 lo = *reinterpret_cast<const double2*>(p);
  const double2 hi = *reinterpret_cast<const double2*>(p + 2);
  o[0] = lo.x; o[1] = lo.y; o[2] = hi.x; o[3] = hi.y;
}

// sum v over the 16 lanes of a rollout IN INDEX ORDER (as the reference's Python loop does),
// result valid in every lane.  `base` = first lane of the group within the warp.
__device__ __forceinline__ double ordered_group_sum(double v, int count, int base, unsigned gmask, double acc) {
  for (int i = 0; i < kLanes; ++i) {
    const double x = __shfl_sync(gmask, v, base + i);
    if (i < count) acc = dadd(acc, x);
  }
  return acc;
}

// Ground truth of the prompts a CTA touches is staged in shared memory (coalesced copy, one
// barrier): the gating / IoU loops then chase smem (30 cycles) instead of L2 (300 cycles).
// Per prompt: kf_time[K] f64 | gt_box[K*O*Gb*4] f64 | n_obj[K] i32 | n_gtbox[K*O] i32.
// (kf_time is padded to an even count and the record to 16 bytes so that the 16-byte box loads stay aligned)
__host__ __device__ inline int gt_kpad(const o3v_rewards_soa& s) { return (s.K + 1) & ~1; }
__host__ __device__ inline int gt_doubles(const o3v_rewards_soa& s) { return gt_kpad(s) + s.K * s.O * s.Gb * 4; }
__host__ __device__ inline int gt_ints(const o3v_rewards_soa& s) { return s.K + s.K * s.O; }
__host__ __device__ inline size_t gt_bytes_per_prompt(const o3v_rewards_soa& s) {
  return (size_t)gt_doubles(s) * 8 + (((size_t)gt_ints(s) * 4 + 15) & ~(size_t)15);
}

template <bool kStageGT, int kRewardThreads>
__global__ void __launch_bounds__(kRewardThreads, 768 / kRewardThreads)      // 24 resident warps per SM
rewards_kernel(const o3v_rewards_soa s, double* __restrict__ out) {
  extern __shared__ double smem_gt[];
  const int lane = threadIdx.x & (kLanes - 1);
  const int64_t r0 = (int64_t)blockIdx.x * (kRewardThreads / kLanes);
  const int64_t r_raw = r0 + (threadIdx.x / kLanes);
  const bool live = r_raw < s.R;
  const int64_t r = live ? r_raw : s.R - 1;      // a group past the end only takes part in the staging barrier
  const int base = (threadIdx.x & 31) & ~(kLanes - 1);
  const unsigned gmask = 0xffffu << base;
  const int64_t q_first = r0 / s.G;
  const int64_t q = r / s.G;

  // ---- load round 1: every scalar of the rollout and of its prompt (no load depends on another one)
  const int flags = s.flags[r];
  const int task = s.task[q];
  const int n_times = s.n_times[r];
  const int nc = s.n_claims[r];
  const int ntb = s.n_tboxes[r];
  const unsigned tvalid = s.tbox_valid[r];
  const int nk = s.n_kf[q];
  const int gtf = s.gt_flags[q];
  const double sp = s.step_percent[q];
  const double as0 = s.ans_seg[r * 2], as1 = s.ans_seg[r * 2 + 1];
  const double gs0 = s.gt_seg[q * 2], gs1 = s.gt_seg[q * 2 + 1];
  const double W = s.image_size[q * 2], H = s.image_size[q * 2 + 1];
  const double rw = s.image_refine[q * 2], rh = s.image_refine[q * 2 + 1];
  double vraw[4], abox[4];
  load4(s.gt_vbox + q * 4, vraw);
  load4(s.ans_box + r * 4, abox);

  if constexpr (kStageGT) {
    const int64_t r_last = min(r0 + kRewardThreads / kLanes, s.R) - 1;
    const int nq = (int)(r_last / s.G - q_first) + 1;
    const int nd = gt_doubles(s), ni = gt_ints(s);
    const size_t stride = gt_bytes_per_prompt(s);
    for (int i = threadIdx.x; i < nq * nd; i += kRewardThreads) {
      const int pq = i / nd, j = i - pq * nd;
      const int64_t qq = q_first + pq;
      double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(smem_gt) + pq * stride);
      const int kp = gt_kpad(s);
      dst[j] = (j < kp) ? (j < s.K ? s.kf_time[qq * s.K + j] : 0.0) : s.gt_box[qq * (int64_t)(nd - kp) + (j - kp)];
    }
    for (int i = threadIdx.x; i < nq * ni; i += kRewardThreads) {
      const int pq = i / ni, j = i - pq * ni;
      const int64_t qq = q_first + pq;
      int32_t* dst = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(smem_gt) + pq * stride + (size_t)nd * 8);
      dst[j] = (j < s.K) ? s.n_obj[qq * s.K + j] : s.n_gtbox[qq * (int64_t)(ni - s.K) + (j - s.K)];
    }
  }

  const bool has_think = flags & O3V_RF_HAS_THINK;
  const bool has_answer = flags & O3V_RF_HAS_ANSWER;
  const bool general = (task == O3V_TASK_GENERAL_MCQ || task == O3V_TASK_GENERAL_FREEFORM);
  const bool temporal = (task == O3V_TASK_TEMPORAL_QA || task == O3V_TASK_TEMPORAL_QA_MCQ);
  const bool visual = (task == O3V_TASK_VISUAL_QA);
  const bool do_seg = has_think && !(visual || task == O3V_TASK_TS_FREEFORM || general);      // reward_func.py:396
  const bool do_point = has_think && !(visual || temporal || general) && n_times > 0;         // :439
  const bool do_vthink = has_think && has_answer && visual;                                    // :490
  const bool do_claims = has_think && has_answer && !visual && !(temporal || general);         // :528

  // ---- load round 2: this lane's timestamp / claim / boxes of the first 16-wide batch, predicated on round 1
  double t_mine = 0.0;
  if ((do_seg || do_point) && lane < n_times) t_mine = s.think_times[r * s.P + lane];
  double ct_mine = 0.0;
  int cnb_mine = 0;
  unsigned cval_mine = 0u;
  double cb0_mine[4] = {0, 0, 0, 0}, cb1_mine[4] = {0, 0, 0, 0};   // claim boxes 0 / 1, or (visual QA) this lane's think box
  if (do_claims && lane < nc) {
    ct_mine = s.claim_t[r * s.C + lane];
    cnb_mine = s.claim_nbox[r * s.C + lane];
    cval_mine = s.claim_valid[r * s.C + lane];
    const double* cb = s.claim_box + ((r * s.C + lane) * (int64_t)s.Bc) * 4;   // slots exist up to Bc whatever the count
    if (s.Bc > 0) load4(cb, cb0_mine);
    if (s.Bc > 1) load4(cb + 4, cb1_mine);
  } else if (do_vthink && lane < ntb) {
    load4(s.think_box + (r * s.Tb + lane) * 4, cb0_mine);
  }

  if constexpr (kStageGT) __syncthreads();
  if (!live) return;   // whole 16-lane group leaves together

  // per-prompt GT arrays: shared-memory copies when staged, else the global arrays
  const double* kft; const double* gtb; const int32_t* nobj_p; const int32_t* ngt_p;
  if constexpr (kStageGT) {
    const char* rec = reinterpret_cast<const char*>(smem_gt) + (size_t)(q - q_first) * gt_bytes_per_prompt(s);
    kft = reinterpret_cast<const double*>(rec);
    gtb = kft + gt_kpad(s);
    nobj_p = reinterpret_cast<const int32_t*>(rec + (size_t)gt_doubles(s) * 8);
    ngt_p = nobj_p + s.K;
  } else {
    kft = s.kf_time + q * s.K;
    gtb = s.gt_box + q * (int64_t)s.K * s.O * s.Gb * 4;
    nobj_p = s.n_obj + q * s.K;
    ngt_p = s.n_gtbox + q * (int64_t)s.K * s.O;
  }

  double r_tiou = 0.0, r_viou = 0.0, r_seg = 0.0, r_point = 0.0, r_spatial = 0.0;

  // ---- ans_tiou_reward (reward_func.py:99-143)
  if (temporal && (flags & O3V_RF_ANS_SEG)) {
    const double s1 = as0, e1 = as1, s2 = gs0, e2 = gs1;
    if (!(e1 < s1)) {                                                   // :128
      const double inter = fmax(0.0, dsub(fmin(e1, e2), fmax(s1, s2))); // :138-140
      const double uni = dsub(fmax(e1, e2), fmin(s1, s2));              // :141
      r_tiou = (uni != 0.0) ? __ddiv_rn(inter, uni) : 0.0;              // :142
    }
  }

  // GT box of the visual-QA tasks, rescaled (convert_coord_format_gqa, :349-354)
  double gvb[4] = {0, 0, 0, 0};
  const bool has_gvb = visual && (gtf & O3V_GF_VBOX);
  if (has_gvb) {
    gvb[0] = __ddiv_rn(dmul(vraw[0], rw), W); gvb[1] = __ddiv_rn(dmul(vraw[1], rh), H);
    gvb[2] = __ddiv_rn(dmul(vraw[2], rw), W); gvb[3] = __ddiv_rn(dmul(vraw[3], rh), H);
  }

  // ---- ans_viou_reward (:196-226)
  if (has_gvb && (flags & O3V_RF_ANS_BOX)) r_viou = box_iou(gvb, abox);

  // ---- thk_temporal_segment_reward (:396, :416-420)
  if (do_seg) {
    int hits = 0;
    for (int p = lane; p < n_times; p += kLanes) {
      const double t = (p == lane) ? t_mine : s.think_times[r * s.P + p];
      hits += (gs0 <= t && t <= gs1) ? 1 : 0;
    }
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) hits += __shfl_xor_sync(gmask, hits, o);
    if (n_times > 0) r_seg = __ddiv_rn((double)hits, (double)n_times);   // exact small integers
  }

  // ---- thk_temporal_point_reward: adaptive temporal proximity (:439, :452-467)
  if (do_point) {
    const double sigma = (sp < 0.75) ? dmul(4.0, dsub(1.0, sp)) : 1.0;  // :459-462
    const double two_s2 = dmul(2.0, dmul(sigma, sigma));
    double total = 0.0;
    for (int p0 = 0; p0 < n_times; p0 += kLanes) {
      const int p = p0 + lane;
      double score = 0.0;
      if (p < n_times) {
        const double t = (p0 == 0) ? t_mine : s.think_times[r * s.P + p];
        double d = INFINITY;
        for (int k = 0; k < nk; ++k) d = fmin(d, fabs(dsub(t, kft[k])));   // :457
        score = exp(__ddiv_rn(-dmul(d, d), two_s2));                       // :463
      }
      total = ordered_group_sum(score, min(kLanes, n_times - p0), base, gmask, total);
    }
    r_point = __ddiv_rn(total, (double)n_times);                           // :467
  }

  // ---- thk_spatial_reward (:484-603)
  if (do_vthink) {                                                         // :490-525
    if (ntb > 0 && has_gvb) {
      double best = 0.0;
      for (int b = lane; b < ntb; b += kLanes) {
        if (b < 32 ? ((tvalid >> b) & 1u) != 0u : box_slot_valid(s.think_box + (r * s.Tb + b) * 4)) {
          if (b == lane) {
            best = fmax(best, box_iou(gvb, cb0_mine));
          } else {
            double pb[4];
            load4(s.think_box + (r * s.Tb + b) * 4, pb);
            best = fmax(best, box_iou(gvb, pb));
          }
        }
      }
#pragma unroll
      for (int o = kLanes / 2; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(gmask, best, o));
      r_spatial = best;
    }
  } else if (do_claims && nc > 0) {
    double total = 0.0;
    for (int c0 = 0; c0 < nc; c0 += kLanes) {
      const int c = c0 + lane;
      double score = 0.0;
      if (c < nc) {
        const bool pre = (c0 == 0);                                        // first batch: loaded in round 2
        const double t = pre ? ct_mine : s.claim_t[r * s.C + c];
        // temporal gating (:550-560): one-sided signed test, strict '<' keeps the first
        int kbest = -1;
        double dbest = INFINITY, tbest = -1.0;
        for (int k = 0; k < nk; ++k) {
          const double g = kft[k];
          if (dsub(g, t) < 1.0) {
            const double d = fabs(dsub(g, t));
            if (d < dbest) { dbest = d; tbest = g; kbest = k; }
          }
        }
        // :561 sentinel compare on the VALUE (a key frame at exactly -1 s is "none")
        if (kbest >= 0 && tbest != -1.0) {
          // :566-569 first key frame whose time equals the chosen one
          int kf = kbest;
          for (int k = 0; k < nk; ++k) if (kft[k] == tbest) { kf = k; break; }
          const int nb = pre ? cnb_mine : s.claim_nbox[r * s.C + c];
          const unsigned valid = pre ? cval_mine : s.claim_valid[r * s.C + c];
          const double* cb = s.claim_box + ((r * s.C + c) * (int64_t)s.Bc) * 4;
          double cb0[4] = {0, 0, 0, 0}, cb1[4] = {0, 0, 0, 0};          // the common case: <= 2 boxes per claim
          if (pre) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { cb0[j] = cb0_mine[j]; cb1[j] = cb1_mine[j]; }
          } else {
            if (nb > 0) load4(cb, cb0);
            if (nb > 1) load4(cb + 4, cb1);
          }
          const int nobj = nobj_p[kf];
          double max_iou = 0.0;
          for (int o = 0; o < nobj; ++o) {                              // :575
            const int ng = ngt_p[kf * s.O + o];
            if (ng <= 0) continue;                                      // :596 empty list
            double acc = 0.0;
            for (int gi = 0; gi < ng; ++gi) {                           // :590
              double nb4[4], g4[4];
              load4(gtb + (((kf * s.O + o) * s.Gb) + gi) * 4, nb4);
              g4[0] = dmul(nb4[0], W); g4[1] = dmul(nb4[1], H);         // :337-346
              g4[2] = dmul(nb4[2], W); g4[3] = dmul(nb4[3], H);
              // :592-593 max(list): boxes 0 and 1 (the common case) as two INDEPENDENT chains, selected afterwards
              const double v0 = box_iou(g4, cb0), v1 = box_iou(g4, cb1);
              double best = 0.0;
              if (nb > 0) best = (valid & 1u) ? v0 : 0.0;
              if (nb > 1) best = fmax(best, (valid & 2u) ? v1 : 0.0);
              for (int b = 2; b < nb; ++b) {
                double v = 0.0;
                if (b < 32 ? ((valid >> b) & 1u) != 0u : box_slot_valid(cb + b * 4)) {
                  double pb[4];
                  load4(cb + b * 4, pb);
                  v = box_iou(g4, pb);
                }
                best = fmax(best, v);
              }
              acc = dadd(acc, best);                                    // :597 sum(...)
            }
            const double iou = __ddiv_rn(acc, (double)ng);
            if (iou > max_iou) max_iou = iou;                           // :598-599
          }
          score = max_iou;
        }
      }
      total = ordered_group_sum(score, min(kLanes, nc - c0), base, gmask, total);   // :601
    }
    r_spatial = __ddiv_rn(total, (double)nc);                           // :603
  }

  if (lane == 0) {
    double* o = out + r * 5;
    o[0] = r_tiou; o[1] = r_viou; o[2] = r_seg; o[3] = r_point; o[4] = r_spatial;
  }
}

}  // namespace o3v

extern "C" int o3v_grounded_rewards(const o3v_rewards_soa* soa, double* out, void* stream) {
  if (!soa || !out) return O3V_ERR_INVALID_ARG;
  const o3v_rewards_soa& s = *soa;
  if (s.R < 0 || s.G <= 0 || (s.R % s.G) != 0) return O3V_ERR_INVALID_ARG;
  if (s.P < 0 || s.C < 0 || s.Bc < 0 || s.Tb < 0 || s.K < 0 || s.O < 0 || s.Gb < 0)
    return O3V_ERR_INVALID_ARG;
  if (!s.flags || !s.ans_seg || !s.ans_box || !s.n_times || !s.think_times || !s.n_claims || !s.claim_t ||
      !s.claim_nbox || !s.claim_valid || !s.claim_box || !s.n_tboxes || !s.tbox_valid || !s.think_box ||
      !s.task || !s.step_percent || !s.gt_flags || !s.gt_seg || !s.gt_vbox || !s.image_size || !s.image_refine || !s.n_kf ||
      !s.kf_time || !s.n_obj || !s.n_gtbox || !s.gt_box)
    return O3V_ERR_INVALID_ARG;
  const uintptr_t al = (uintptr_t)s.ans_box | (uintptr_t)s.claim_box | (uintptr_t)s.think_box |
                       (uintptr_t)s.gt_vbox | (uintptr_t)s.gt_box;
  if (al & 15u) return O3V_ERR_ALIGNMENT;
  int rc = o3v::check_device();
  if (rc) return rc;
  if (s.R == 0) return O3V_OK;
  // threads per CTA (diagnostic override: O3V_REWARDS_CTA=64|128|256)
  static const int cta = [] {
    const char* e = getenv("O3V_REWARDS_CTA");
    const int v = e ? atoi(e) : o3v::kRewardThreadsDefault;
    return (v == 64 || v == 128 || v == 256) ? v : o3v::kRewardThreadsDefault;
  }();
  const int per_cta = cta / o3v::kLanes;
  const unsigned grid = (unsigned)((s.R + per_cta - 1) / per_cta);
  // prompts a CTA of `per_cta` consecutive rollouts can touch
  const int64_t span = std::min<int64_t>(per_cta, (per_cta + s.G - 1) / s.G + 1);
  const size_t smem = (size_t)span * o3v::gt_bytes_per_prompt(s);
  cudaStream_t st = (cudaStream_t)stream;
  const bool stage = smem <= (size_t)(40 * 1024) * cta / 256;     // very large K x O x Gb: read the ground truth through L2
#define O3V_REWARDS_LAUNCH(STAGE, THREADS) \
  o3v::rewards_kernel<STAGE, THREADS><<<grid, THREADS, (STAGE) ? smem : 0, st>>>(s, out)
  if (stage) {
    if (cta == 64) O3V_REWARDS_LAUNCH(true, 64); else if (cta == 128) O3V_REWARDS_LAUNCH(true, 128); else O3V_REWARDS_LAUNCH(true, 256);
  } else {
    if (cta == 64) O3V_REWARDS_LAUNCH(false, 64); else if (cta == 128) O3V_REWARDS_LAUNCH(false, 128); else O3V_REWARDS_LAUNCH(false, 256);
  }
#undef O3V_REWARDS_LAUNCH
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
