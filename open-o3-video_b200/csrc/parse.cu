// K6: completion text -> rollout side of o3v_rewards_soa, on the GPU.
//
// Three launches over UTF-8 bytes that are already in HBM (scan_core.cuh has the scanner and the
// reference lines it replaces):
//   A  parse_scan_kernel     one warp per rollout: warp-cooperative literal searches resolve the
//                            <think>/<answer> spans and the regex match chains -> candidate ranges
//   B  parse_convert_kernel  one thread per candidate: decimal -> binary64, JSON box payloads
//   C  parse_finish_kernel   one thread per rollout: drop rejected candidates, compact, counts
// HBM-bound byte work in principle (the text is read once from DRAM in A; B re-reads the few
// bytes of each candidate out of L2); in practice bound by dependent byte loads of the sequential
// number / JSON routines, which is why those run one candidate per THREAD instead of per warp.
#include "common.cuh"
#include "scan_core.cuh"

namespace o3v {

constexpr int kScanWarps = 8;        // phase A: rollouts per CTA
constexpr int kConvertThreads = 128; // phase B
constexpr int kFinishThreads = 128;  // phase C

__device__ __forceinline__ scan::RolloutOut rows_of(const o3v_parse_args& a, int64_t r) {
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
  return o;
}

__global__ void __launch_bounds__(kScanWarps * 32, 4)
parse_scan_kernel(const o3v_parse_args a, scan::Scratch* __restrict__ scratch) {
  extern __shared__ uint32_t mask_cache[];   // kScanWarps x Finder::kSmemWords (6 KB + a 32-entry list per warp)
  const int64_t r = (int64_t)blockIdx.x * kScanWarps + (threadIdx.x >> 5);
  if (r >= a.R) return;                      // whole warp leaves together
  const scan::Caps cap{a.P, a.C, a.Bc, a.Tb};
  scan::scan_rollout(a.text, a.offsets[a.R], a.offsets[r], a.offsets[r + 1], a.task[r / a.G], cap, rows_of(a, r), scratch + r,
                     mask_cache + (threadIdx.x >> 5) * scan::Finder::kSmemWords);
}

__global__ void __launch_bounds__(kConvertThreads, 4)
parse_convert_kernel(const o3v_parse_args a, scan::Scratch* __restrict__ scratch) {
  const scan::Caps cap{a.P, a.C, a.Bc, a.Tb};
  // item-major: the 128 threads of a CTA convert the SAME item of 128 consecutive rollouts, so every lane
  // of a warp runs the same routine (all time candidates k, or all claims c, ...)
  const int per = scan::items_per_rollout(cap);
  const int64_t tiles = (a.R + kConvertThreads - 1) / kConvertThreads;
  const int64_t tile = blockIdx.x % tiles;
  const int item = (int)(blockIdx.x / tiles);
  const int64_t r = tile * kConvertThreads + threadIdx.x;
  const bool in_range = r < a.R && item < per;
  const scan::RolloutOut o = rows_of(a, in_range ? r : 0);
  const bool active = in_range && scan::item_active(item, cap, o, scratch + r);
  const unsigned lanes = __ballot_sync(0xffffffffu, active);   // the lanes that convert: they stay in lockstep
  if (!active) return;
  scan::convert_item(a.text, item, cap, o, scratch + r, lanes);
}

__global__ void __launch_bounds__(kFinishThreads)
parse_finish_kernel(const o3v_parse_args a, const scan::Scratch* __restrict__ scratch) {
  const int64_t r = (int64_t)blockIdx.x * kFinishThreads + threadIdx.x;
  if (r >= a.R) return;
  const scan::Caps cap{a.P, a.C, a.Bc, a.Tb};
  int over[4];
  scan::finish_rollout(cap, rows_of(a, r), scratch + r, over);
#pragma unroll
  for (int i = 0; i < 4; ++i)                // rare: only counts that did not fit are published
    if (over[i]) atomicMax(a.overflow + i, over[i]);
}

}  // namespace o3v

extern "C" size_t o3v_parse_workspace_bytes(int64_t R) {
  return (size_t)(R > 0 ? R : 0) * sizeof(o3v::scan::Scratch);
}

extern "C" int o3v_parse_completions(const o3v_parse_args* args, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  static_assert(sizeof(o3v::scan::Scratch) == 40, "scratch record layout");
  if (!args) return O3V_ERR_INVALID_ARG;
  const o3v_parse_args& a = *args;
  if (a.R < 0 || a.G <= 0 || (a.R % a.G) != 0) return O3V_ERR_INVALID_ARG;
  if (a.P < 1 || a.C < 1 || a.Bc < 1 || a.Bc > 32 || a.Tb < 1 || a.Tb > 32) return O3V_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.R == 0) {
    if (a.overflow) O3V_CUDA_TRY(cudaMemsetAsync(a.overflow, 0, 4 * sizeof(int32_t), st));
    return O3V_OK;
  }
  if (!a.text || !a.offsets || !a.task || !a.flags || !a.ans_seg || !a.ans_box || !a.n_times || !a.think_times ||
      !a.n_claims || !a.claim_t || !a.claim_nbox || !a.claim_valid || !a.claim_box || !a.n_tboxes ||
      !a.tbox_valid || !a.think_box || !a.overflow)
    return O3V_ERR_INVALID_ARG;
  if (((uintptr_t)a.text & 15u) || ((uintptr_t)workspace & 7u)) return O3V_ERR_ALIGNMENT;
  if (!workspace || workspace_bytes < o3v_parse_workspace_bytes(a.R)) return O3V_ERR_WORKSPACE;
  int rc = o3v::check_device();
  if (rc) return rc;
  auto* scratch = reinterpret_cast<o3v::scan::Scratch*>(workspace);
  O3V_CUDA_TRY(cudaMemsetAsync(a.overflow, 0, 4 * sizeof(int32_t), st));
  const unsigned grid_a = (unsigned)((a.R + o3v::kScanWarps - 1) / o3v::kScanWarps);
  const size_t smem_a = (size_t)o3v::kScanWarps * o3v::scan::Finder::kSmemWords * sizeof(uint32_t);
  O3V_CUDA_TRY(cudaFuncSetAttribute(o3v::parse_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  o3v::parse_scan_kernel<<<grid_a, o3v::kScanWarps * 32, smem_a, st>>>(a, scratch);
  O3V_LAUNCH_CHECK();
  const int64_t tiles = (a.R + o3v::kConvertThreads - 1) / o3v::kConvertThreads;
  const unsigned grid_b = (unsigned)(tiles * (a.P + a.C + a.Tb + 2));
  o3v::parse_convert_kernel<<<grid_b, o3v::kConvertThreads, 0, st>>>(a, scratch);
  O3V_LAUNCH_CHECK();
  const unsigned grid_c = (unsigned)((a.R + o3v::kFinishThreads - 1) / o3v::kFinishThreads);
  o3v::parse_finish_kernel<<<grid_c, o3v::kFinishThreads, 0, st>>>(a, scratch);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
