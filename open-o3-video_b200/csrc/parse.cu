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

// Work lists between A and B: A appends every candidate of the three list-like item types to a dense
// list (one atomicAdd per rollout and type reserves the range), so that B's warps are full however
// unevenly the candidates are spread over the rollouts.  Entry = rollout * capacity + slot.
struct WorkLists {
  unsigned int* count;    // [4]: claims, think times, think boxes, (unused)
  uint32_t* claims;       // [R * C]
  uint32_t* times;        // [R * P]
  uint32_t* tboxes;       // [R * Tb]
};
__host__ __device__ inline size_t lists_offset(int64_t R) {
  return ((size_t)R * sizeof(scan::Scratch) + 15) & ~(size_t)15;
}
__host__ __device__ inline WorkLists work_lists(void* workspace, const o3v_parse_args& a) {
  char* base = reinterpret_cast<char*>(workspace) + lists_offset(a.R);
  WorkLists w;
  w.count = reinterpret_cast<unsigned int*>(base);
  w.claims = reinterpret_cast<uint32_t*>(base + 16);
  w.times = w.claims + (size_t)a.R * a.C;
  w.tboxes = w.times + (size_t)a.R * a.P;
  return w;
}

// lanes 0..2 reserve the three ranges with one atomicAdd each (in flight together), then all lanes fill them
__device__ __forceinline__ void append_candidates(const WorkLists& w, const o3v_parse_args& a, int64_t r,
                                                  int n_claims, int n_times, int n_tboxes) {
  const int lane = threadIdx.x & 31;
  const int mine = lane == 0 ? n_claims : lane == 1 ? n_times : lane == 2 ? n_tboxes : 0;
  unsigned int base = 0;
  if (mine > 0) base = atomicAdd(w.count + lane, (unsigned int)mine);
  const unsigned int b0 = __shfl_sync(0xffffffffu, base, 0), b1 = __shfl_sync(0xffffffffu, base, 1),
                     b2 = __shfl_sync(0xffffffffu, base, 2);
  for (int k = lane; k < n_claims; k += 32) w.claims[b0 + k] = (uint32_t)(r * a.C) + (uint32_t)k;
  for (int k = lane; k < n_times; k += 32) w.times[b1 + k] = (uint32_t)(r * a.P) + (uint32_t)k;
  for (int k = lane; k < n_tboxes; k += 32) w.tboxes[b2 + k] = (uint32_t)(r * a.Tb) + (uint32_t)k;
}

__global__ void __launch_bounds__(kScanWarps * 32, 4)
parse_scan_kernel(const o3v_parse_args a, scan::Scratch* __restrict__ scratch, const WorkLists w) {
  extern __shared__ uint32_t mask_cache[];   // kScanWarps x Finder::kSmemWords (6 KB + a 32-entry list per warp)
  const int64_t r = (int64_t)blockIdx.x * kScanWarps + (threadIdx.x >> 5);
  if (r >= a.R) return;                      // whole warp leaves together
  const scan::Caps cap{a.P, a.C, a.Bc, a.Tb};
  scan::Scratch* sc = scratch + r;
  scan::scan_rollout(a.text, a.offsets[a.R], a.offsets[r], a.offsets[r + 1], a.task[r / a.G], cap, rows_of(a, r), sc,
                     mask_cache + (threadIdx.x >> 5) * scan::Finder::kSmemWords);
  __syncwarp();                              // lane 0 wrote the counts
  append_candidates(w, a, r, min(sc->claim_cands, a.C), min(sc->time_cands, a.P), min(sc->tbox_cands, a.Tb));
}

// B: CTA ranges [claims | think times | think boxes | answers]; within a range thread i converts list entry i
// (answers: rollout i / 2, segment or box), so all lanes of a warp run the same routine on dense work.
__global__ void __launch_bounds__(kConvertThreads, 4)
parse_convert_kernel(const o3v_parse_args a, scan::Scratch* __restrict__ scratch, const WorkLists w,
                     unsigned nb_claims, unsigned nb_times, unsigned nb_tboxes) {
  const scan::Caps cap{a.P, a.C, a.Bc, a.Tb};
  unsigned b = blockIdx.x;
  int64_t r = 0;
  int item = 0;
  bool active = false;
  if (b < nb_claims) {
    const int64_t i = (int64_t)b * kConvertThreads + threadIdx.x;
    if (i < (int64_t)w.count[0]) { const uint32_t e = w.claims[i]; r = e / (uint32_t)a.C; item = a.P + (int)(e % (uint32_t)a.C); active = true; }
  } else if ((b -= nb_claims) < nb_times) {
    const int64_t i = (int64_t)b * kConvertThreads + threadIdx.x;
    if (i < (int64_t)w.count[1]) { const uint32_t e = w.times[i]; r = e / (uint32_t)a.P; item = (int)(e % (uint32_t)a.P); active = true; }
  } else if ((b -= nb_times) < nb_tboxes) {
    const int64_t i = (int64_t)b * kConvertThreads + threadIdx.x;
    if (i < (int64_t)w.count[2]) { const uint32_t e = w.tboxes[i]; r = e / (uint32_t)a.Tb; item = a.P + a.C + (int)(e % (uint32_t)a.Tb); active = true; }
  } else {
    b -= nb_tboxes;                                                     // answers: a CTA range of segments, then one of boxes
    const unsigned nb_rollouts = (unsigned)((a.R + kConvertThreads - 1) / kConvertThreads);
    const int which = b >= nb_rollouts ? 1 : 0;                         // (a warp never mixes item types: its lanes share
    r = (int64_t)(b - which * nb_rollouts) * kConvertThreads + threadIdx.x;   //  lockstep loops)
    item = a.P + a.C + a.Tb + which;
    active = r < a.R && scan::item_active(item, cap, rows_of(a, r), scratch + r);
  }
  const unsigned lanes = __ballot_sync(0xffffffffu, active);   // the lanes that convert: they stay in lockstep
  if (!active) return;
  scan::convert_item(a.text, item, cap, rows_of(a, r), scratch + r, lanes);
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

extern "C" size_t o3v_parse_workspace_bytes(int64_t R, int32_t P, int32_t C, int32_t Tb) {
  if (R <= 0 || P < 0 || C < 0 || Tb < 0) return 0;
  return o3v::lists_offset(R) + 16 + (size_t)R * ((size_t)P + C + Tb) * sizeof(uint32_t);
}

extern "C" int o3v_parse_completions(const o3v_parse_args* args, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  static_assert(sizeof(o3v::scan::Scratch) == 40, "scratch record layout");
  if (!args) return O3V_ERR_INVALID_ARG;
  const o3v_parse_args& a = *args;
  if (a.R < 0 || a.G <= 0 || (a.R % a.G) != 0) return O3V_ERR_INVALID_ARG;
  if (a.P < 1 || a.C < 1 || a.Bc < 1 || a.Tb < 1) return O3V_ERR_INVALID_ARG;
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
  if (!workspace || workspace_bytes < o3v_parse_workspace_bytes(a.R, a.P, a.C, a.Tb)) return O3V_ERR_WORKSPACE;
  const int64_t cap_max = a.P > a.C ? (a.P > a.Tb ? a.P : a.Tb) : (a.C > a.Tb ? a.C : a.Tb);
  if (a.R * cap_max >= ((int64_t)1 << 32)) return O3V_ERR_SHAPE;            // work-list entries are 32-bit
  int rc = o3v::check_device();
  if (rc) return rc;
  auto* scratch = reinterpret_cast<o3v::scan::Scratch*>(workspace);
  O3V_CUDA_TRY(cudaMemsetAsync(a.overflow, 0, 4 * sizeof(int32_t), st));
  const unsigned grid_a = (unsigned)((a.R + o3v::kScanWarps - 1) / o3v::kScanWarps);
  const size_t smem_a = (size_t)o3v::kScanWarps * o3v::scan::Finder::kSmemWords * sizeof(uint32_t);
  O3V_CUDA_TRY(cudaFuncSetAttribute(o3v::parse_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  const o3v::WorkLists lists = o3v::work_lists(workspace, a);
  O3V_CUDA_TRY(cudaMemsetAsync(lists.count, 0, 16, st));
  o3v::parse_scan_kernel<<<grid_a, o3v::kScanWarps * 32, smem_a, st>>>(a, scratch, lists);
  O3V_LAUNCH_CHECK();
  auto blocks = [](int64_t n) { return (unsigned)((n + o3v::kConvertThreads - 1) / o3v::kConvertThreads); };
  const unsigned nb_claims = blocks(a.R * a.C), nb_times = blocks(a.R * a.P), nb_tboxes = blocks(a.R * a.Tb);
  const unsigned grid_b = nb_claims + nb_times + nb_tboxes + 2 * blocks(a.R);
  o3v::parse_convert_kernel<<<grid_b, o3v::kConvertThreads, 0, st>>>(a, scratch, lists, nb_claims, nb_times, nb_tboxes);
  O3V_LAUNCH_CHECK();
  const unsigned grid_c = (unsigned)((a.R + o3v::kFinishThreads - 1) / o3v::kFinishThreads);
  o3v::parse_finish_kernel<<<grid_c, o3v::kFinishThreads, 0, st>>>(a, scratch);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
