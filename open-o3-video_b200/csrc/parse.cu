// K6: completion text -> rollout side of o3v_rewards_soa, on the GPU.
//
// One thread per rollout over UTF-8 bytes that are already in HBM; see scan_core.cuh for the
// scanner and for the reference lines it replaces.  HBM-bound byte work: the text is read once
// from DRAM (every later touch of a span hits L1), outputs are written only up to the counts
// found.  The lanes of a warp walk 32 different completions, so loads are per-lane 16-byte
// vectors out of 32 different cache lines; the lines stay in L1 until their thread has consumed
// them (a few hundred bytes of window per thread).
#include "common.cuh"
#include "scan_core.cuh"

namespace o3v {

constexpr int kParseThreads = 128;

__global__ void __launch_bounds__(kParseThreads, 4)
parse_kernel(const o3v_parse_args a) {
  const int64_t r = (int64_t)blockIdx.x * kParseThreads + threadIdx.x;
  if (r >= a.R) return;
  const int64_t beg = a.offsets[r], end = a.offsets[r + 1];
  const int task = a.task[r / a.G];
  scan::Caps cap{a.P, a.C, a.Bc, a.Tb};
  scan::RolloutOut o;
  o.flags = a.flags + r;
  o.ans_seg = a.ans_seg + r * 2;
  o.ans_box = a.ans_box + r * 4;
  o.n_times = a.n_times + r;
  o.think_times = a.think_times + r * a.P;
  o.n_claims = a.n_claims + r;
  o.claim_t = a.claim_t + r * a.C;
  o.claim_nbox = a.claim_nbox + r * a.C;
  o.claim_valid = a.claim_valid + r * a.C;
  o.claim_box = a.claim_box + r * (int64_t)a.C * a.Bc * 4;
  o.n_tboxes = a.n_tboxes + r;
  o.tbox_valid = a.tbox_valid + r;
  o.think_box = a.think_box + r * (int64_t)a.Tb * 4;
  scan::Maxima mx;
  scan::parse_rollout(a.text, beg, end, task, cap, o, &mx);
  // rare: only counts that did not fit are published
  if (mx.times > a.P) atomicMax(a.overflow + 0, mx.times);
  if (mx.claims > a.C) atomicMax(a.overflow + 1, mx.claims);
  if (mx.claim_boxes > a.Bc) atomicMax(a.overflow + 2, mx.claim_boxes);
  if (mx.think_boxes > a.Tb) atomicMax(a.overflow + 3, mx.think_boxes);
}

}  // namespace o3v

extern "C" int o3v_parse_completions(const o3v_parse_args* args, void* stream) {
  if (!args) return O3V_ERR_INVALID_ARG;
  const o3v_parse_args& a = *args;
  if (a.R < 0 || a.G <= 0 || (a.R % a.G) != 0) return O3V_ERR_INVALID_ARG;
  if (a.P < 0 || a.C < 0 || a.Bc < 0 || a.Bc > 32 || a.Tb < 0 || a.Tb > 32) return O3V_ERR_INVALID_ARG;
  if (a.R == 0) {
    if (a.overflow) O3V_CUDA_TRY(cudaMemsetAsync(a.overflow, 0, 4 * sizeof(int32_t), (cudaStream_t)stream));
    return O3V_OK;
  }
  if (!a.text || !a.offsets || !a.task || !a.flags || !a.ans_seg || !a.ans_box || !a.n_times || !a.think_times ||
      !a.n_claims || !a.claim_t || !a.claim_nbox || !a.claim_valid || !a.claim_box || !a.n_tboxes ||
      !a.tbox_valid || !a.think_box || !a.overflow)
    return O3V_ERR_INVALID_ARG;
  if ((uintptr_t)a.text & 15u) return O3V_ERR_ALIGNMENT;
  int rc = o3v::check_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  O3V_CUDA_TRY(cudaMemsetAsync(a.overflow, 0, 4 * sizeof(int32_t), st));
  const unsigned grid = (unsigned)((a.R + o3v::kParseThreads - 1) / o3v::kParseThreads);
  o3v::parse_kernel<<<grid, o3v::kParseThreads, 0, st>>>(a);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
