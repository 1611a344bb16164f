// K6: completion text -> rollout side of o3v_rewards_soa, on the GPU.
//
// One warp per rollout over UTF-8 bytes that are already in HBM; see scan_core.cuh for the
// scanner and for the reference lines it replaces.  HBM-bound byte work: the text is read once
// from DRAM (later passes over the same span hit L1 / L2), outputs are written only up to the
// counts found.  Rollouts are handed out through an atomic ticket so that long completions do not
// leave a CTA's other warps idle.
#include "common.cuh"
#include "scan_core.cuh"

namespace o3v {

constexpr int kParseWarps = 8;   // warps (= rollouts in flight) per CTA

__global__ void __launch_bounds__(kParseWarps * 32, 3)
parse_kernel(const o3v_parse_args a, unsigned long long* __restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long r = 0;
    if (lane == 0) r = atomicAdd(ticket, 1ull);
    r = __shfl_sync(0xffffffffu, r, 0);
    if ((int64_t)r >= a.R) return;
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
    if (lane == 0) {   // rare: only counts that did not fit are published
      if (mx.times > a.P) atomicMax(a.overflow + 0, mx.times);
      if (mx.claims > a.C) atomicMax(a.overflow + 1, mx.claims);
      if (mx.claim_boxes > a.Bc) atomicMax(a.overflow + 2, mx.claim_boxes);
      if (mx.think_boxes > a.Tb) atomicMax(a.overflow + 3, mx.think_boxes);
    }
  }
}

}  // namespace o3v

extern "C" size_t o3v_parse_workspace_bytes(void) { return 16; }

extern "C" int o3v_parse_completions(const o3v_parse_args* args, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (!args) return O3V_ERR_INVALID_ARG;
  const o3v_parse_args& a = *args;
  if (a.R < 0 || a.G <= 0 || (a.R % a.G) != 0) return O3V_ERR_INVALID_ARG;
  if (a.P < 0 || a.C < 0 || a.Bc < 0 || a.Bc > 32 || a.Tb < 0 || a.Tb > 32) return O3V_ERR_INVALID_ARG;
  if (!a.text || !a.offsets || !a.task || !a.flags || !a.ans_seg || !a.ans_box || !a.n_times || !a.think_times ||
      !a.n_claims || !a.claim_t || !a.claim_nbox || !a.claim_valid || !a.claim_box || !a.n_tboxes ||
      !a.tbox_valid || !a.think_box || !a.overflow)
    return O3V_ERR_INVALID_ARG;
  if (((uintptr_t)a.text & 15u) || ((uintptr_t)workspace & 7u)) return O3V_ERR_ALIGNMENT;
  if (!workspace || workspace_bytes < o3v_parse_workspace_bytes()) return O3V_ERR_WORKSPACE;
  int rc = o3v::check_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  O3V_CUDA_TRY(cudaMemsetAsync(a.overflow, 0, 4 * sizeof(int32_t), st));
  if (a.R == 0) return O3V_OK;
  O3V_CUDA_TRY(cudaMemsetAsync(workspace, 0, 8, st));
  // persistent grid: enough CTAs to fill every SM, rollouts drawn from the ticket
  const int64_t want = (a.R + o3v::kParseWarps - 1) / o3v::kParseWarps;
  const int64_t cap = (int64_t)o3v::num_sms() * 3;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  o3v::parse_kernel<<<grid, o3v::kParseWarps * 32, 0, st>>>(a, reinterpret_cast<unsigned long long*>(workspace));
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
