// Host side of K1 / K2 (C ABI in include/o3v.h) + the small elementwise kernels around the
// tcgen05 GEMMs: partial-statistics merge and the in-place softmax backward (dlogits).
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>

#include "lmhead_gemm.cuh"

namespace o3v {

// ------------------------------------------------------------------------------------
// Tunables (diagnostics / bench sweeps; defaults are what the product uses)
// ------------------------------------------------------------------------------------
// tile mode per kernel: 1 = one CTA per 128-row tile (UMMA M=128), 2 = cta_group::2 pair per 256-row tile
static int g_cta_fwd = 1;       // K1 (measured: the 4-stage 1-CTA pipeline is ahead for the K-major sweep)
static int g_cta_bwd = 2;       // K2a / K2b
static int g_bwd_wide = 1;      // K2a / K2b on pairs: 256x512 tiles (both TMEM accumulators per tile)
static int g_bwd_sync = 1;      // K2a / K2b: wave-boundary rendezvous of the TMA producers
static int g_dh_mfast = 0;      // K2a item order (0 = n fastest)
static int g_dw_mfast = 0;      // K2b item order
static int g_fwd_groups = 0;    // n-groups (vocab splits) per m-block in K1; 0 = auto
static int g_max_ctas = 0;      // cap on the persistent grid; 0 = all SMs
static int g_bwd_prefetch = 0;  // K2a / K2b: L2 prefetch distance (k-blocks) for the A operand streamed from HBM
// L2 eviction hints (0 none, 1 evict first, 2 evict last): K1 hidden / W loads and logits store, K2 P / other operand
static int g_hint_fwd_a = 0, g_hint_fwd_b = 0, g_hint_fwd_store = 0, g_hint_bwd_a = 0, g_hint_bwd_b = 0;

// ------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched through the runtime: no libcuda link)
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 row-major tensor [outer, inner] with row pitch `ld` elements; box = [box_outer, 64]
// with the 128-byte swizzle; out-of-bounds elements read as zero.
static int make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t inner, int64_t outer, int64_t ld,
                          int box_outer, int box_inner = 64) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return O3V_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 2) & 15)) return O3V_ERR_ALIGNMENT;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? O3V_OK : O3V_ERR_DRIVER;
}

// ------------------------------------------------------------------------------------
// wave-sync counters: static device memory (the library allocates nothing at run time).  A
// launch takes the next slot round-robin and zeroes it on its stream, so GEMMs that overlap on
// different streams do not share a counter.
// ------------------------------------------------------------------------------------
constexpr int kSyncSlots = 64;
__device__ unsigned int g_wave_sync[kSyncSlots];

static int acquire_wave_sync(unsigned int** out, cudaStream_t st) {
  static std::atomic<unsigned int> next{0};
  unsigned int* base = nullptr;
  O3V_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_wave_sync));
  unsigned int* slot = base + (next.fetch_add(1) % kSyncSlots);
  O3V_CUDA_TRY(cudaMemsetAsync(slot, 0, sizeof(unsigned int), st));
  *out = slot;
  return O3V_OK;
}

// ------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------
template <bool kAMN, bool kBMN, int kNCta, int kEpi, int kAcc = 1, bool kXform = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t st) {
  using S = GemmShape<kNCta, kEpi == EPI_STATS, kAcc>;
  auto kern = lmhead_gemm_kernel<kAMN, kBMN, kNCta, kEpi, kAcc, kXform>;
  // the opt-in shared-memory size is a per-device (per-context) function attribute
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  O3V_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    O3V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  int sms = num_sms();
  if (g_max_ctas > 0 && g_max_ctas < sms) sms = g_max_ctas;
  const int64_t items = (int64_t)p.num_m_blocks * p.num_n_groups;
  int workers = (int)std::min<int64_t>(sms / kNCta, items);
  if (workers < 1) workers = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(workers * kNCta));
  cfg.blockDim = dim3(kXform ? kGemmThreadsXform : kGemmThreads);
  cfg.dynamicSmemBytes = S::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kNCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  O3V_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p));
  return O3V_OK;
}

static void plan_tiles(GemmParams& p, int ncta, int n_groups_hint, int acc = 1) {
  p.num_m_blocks = (int32_t)ceil_div(p.M, 128 * ncta);
  p.num_n_tiles = (int32_t)ceil_div(p.N, 256 * acc);
  int groups = n_groups_hint < 1 ? 1 : n_groups_hint;
  if (groups > p.num_n_tiles) groups = p.num_n_tiles;
  p.tiles_per_group = (int32_t)ceil_div(p.num_n_tiles, groups);
  p.num_n_groups = (int32_t)ceil_div(p.num_n_tiles, p.tiles_per_group);
}

// Number of vocab splits per m-block in K1.  Few splits keep the concurrently swept A panels
// (hidden rows) L2-resident and W re-reads low; enough splits fill the persistent grid.
static int fwd_groups(int64_t T, int64_t V, int ncta) {
  if (g_fwd_groups > 0) return g_fwd_groups;
  const int64_t num_m = ceil_div(T, 128 * ncta);
  const int64_t n_tiles = ceil_div(V, 256);
  const int64_t workers = num_sms() / ncta;
  int64_t g = 4;
  while (num_m * g < 2 * workers && g < n_tiles) g *= 2;     // small T: split the vocab further
  if (g > n_tiles) g = n_tiles;
  return (int)g;
}

// ------------------------------------------------------------------------------------
// merge of partial statistics
// ------------------------------------------------------------------------------------
// parts [P, 3, T] -> either one merged triple [3, T] (triple_out) or (logp, lse).
__global__ void merge_stats_kernel(const float* __restrict__ parts, int64_t P, int64_t T,
                                   float* __restrict__ triple_out, float* __restrict__ logp,
                                   float* __restrict__ lse) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float M = -INFINITY;
  for (int64_t i = 0; i < P; ++i) M = fmaxf(M, parts[(i * 3 + 0) * T + t]);
  float s = 0.f, z = 0.f;
  for (int64_t i = 0; i < P; ++i) {            // fixed order: deterministic
    const float m = parts[(i * 3 + 0) * T + t];
    const float e = (m == -INFINITY) ? 0.f : expf(m - M);
    s += parts[(i * 3 + 1) * T + t] * e;
    z += parts[(i * 3 + 2) * T + t];           // non-zero in the owner slice only
  }
  if (triple_out) {
    triple_out[t] = M;
    triple_out[T + t] = s;
    triple_out[2 * T + t] = z;
  } else {
    const float l = M + logf(s);
    if (lse) lse[t] = l;
    if (logp) logp[t] = z - l;
  }
}

// Same merge, reading each rank's [3, row_stride] triple through its peer-mapped pointer (NVLink
// loads, 12 B per token per rank): the all-gather and the merge are one kernel.
struct PeerPtrs { const float* p[16]; };
__global__ void merge_stats_peers_kernel(const PeerPtrs ptrs, int P, int64_t row_stride, int64_t T,
                                         float* __restrict__ logp, float* __restrict__ lse) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float m[16], sv[16], z = 0.f, M = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i < P) {                                   // issue every peer load before the first use
      m[i] = __ldcv(ptrs.p[i] + t);
      sv[i] = __ldcv(ptrs.p[i] + row_stride + t);
      z += __ldcv(ptrs.p[i] + 2 * row_stride + t);
      M = fmaxf(M, m[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < P) s += sv[i] * ((m[i] == -INFINITY) ? 0.f : expf(m[i] - M));   // fixed rank order: deterministic
  const float l = M + logf(s);
  if (lse) lse[t] = l;
  if (logp) logp[t] = z - l;
}

// One-shot P2P all-reduce of bf16 partial sums (dHidden of the vocab-parallel path).
struct PeerBufs { uint4* p[16]; };
constexpr int kArThreads = 256;
// <= 64 registers x 256 threads and no shared memory: fits beside a persistent GEMM CTA on the same SM
__global__ void __launch_bounds__(kArThreads, 4)
allreduce_bf16_peers_kernel(const PeerBufs bufs, int P, int64_t v0, int64_t v1) {
  // this rank's slice of 16-byte vectors [v0, v1), grid-stride; peers are read four at a time
  const int64_t stride = (int64_t)gridDim.x * kArThreads;
  for (int64_t i = v0 + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i < v1; i += stride) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p0 = 0; p0 < P; p0 += 4) {
      uint4 a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + j < P) a[j] = __ldcv(bufs.p[p0 + j] + i);   // four NVLink loads in flight
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + j < P) {                                    // fixed rank order: identical result everywhere
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(h[k]);
            s[2 * k] += f.x; s[2 * k + 1] += f.y;
          }
        }
    }
    uint4 r;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = __floats2bfloat162_rn(s[2 * k], s[2 * k + 1]);
    for (int p = 0; p < P; ++p) bufs.p[p][i] = r;
  }
}

// ------------------------------------------------------------------------------------
// dlogits: P[t,v] = g[t] * ([v + v_offset == target[t]] - exp(z[t,v] - lse[t])), in place, bf16
// ------------------------------------------------------------------------------------
constexpr int kDlThreads = 256;
__global__ void __launch_bounds__(kDlThreads)
dlogits_kernel(__nv_bfloat16* __restrict__ z, int64_t T, int64_t V, int64_t ld, const float* __restrict__ lse,
               const float* __restrict__ g, const int64_t* __restrict__ targets, int64_t v_offset,
               int rows_per_cta) {
  const int64_t t0 = (int64_t)blockIdx.x * rows_per_cta;
  for (int r = 0; r < rows_per_cta; ++r) {
    const int64_t t = t0 + r;
    if (t >= T) return;
    const float gt = g[t];
    const float neg_l = -lse[t] * kLog2e;
    const int64_t tc = targets[t] - v_offset;
    uint4* row = reinterpret_cast<uint4*>(z + t * ld);
    const int64_t nvec = V >> 3;                       // V % 8 == 0 (checked by the API)
    for (int64_t i = threadIdx.x; i < nvec; i += kDlThreads) {
      uint4 pk = (gt == 0.f) ? make_uint4(0, 0, 0, 0) : row[i];
      if (gt != 0.f) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          float a = -gt * exp2f(fmaf(f.x, kLog2e, neg_l));
          float b = -gt * exp2f(fmaf(f.y, kLog2e, neg_l));
          const int64_t c = i * 8 + j * 2;
          if (c == tc) a += gt;
          if (c + 1 == tc) b += gt;
          h[j] = __floats2bfloat162_rn(a, b);
        }
      }
      row[i] = pk;
    }
  }
}

}  // namespace o3v

using namespace o3v;

extern "C" int o3v_set_tunable(const char* name, int value) {
  if (!name) return O3V_ERR_INVALID_ARG;
  std::string n(name);
  if (n == "cta_pair") { if (value != 1 && value != 2) return O3V_ERR_INVALID_ARG; g_cta_fwd = g_cta_bwd = value; }
  else if (n == "cta_pair_fwd") { if (value != 1 && value != 2) return O3V_ERR_INVALID_ARG; g_cta_fwd = value; }
  else if (n == "cta_pair_bwd") { if (value != 1 && value != 2) return O3V_ERR_INVALID_ARG; g_cta_bwd = value; }
  else if (n == "bwd_wide") g_bwd_wide = value ? 1 : 0;
  else if (n == "bwd_sync") g_bwd_sync = value ? 1 : 0;
  else if (n == "dh_mfast") g_dh_mfast = value ? 1 : 0;
  else if (n == "dw_mfast") g_dw_mfast = value ? 1 : 0;
  else if (n == "fwd_groups") g_fwd_groups = value;
  else if (n == "max_ctas") g_max_ctas = value;
  else if (n == "bwd_prefetch") g_bwd_prefetch = value < 0 ? 0 : value;
  else if (n == "hint_fwd_a") g_hint_fwd_a = value;
  else if (n == "hint_fwd_b") g_hint_fwd_b = value;
  else if (n == "hint_fwd_store") g_hint_fwd_store = value;
  else if (n == "hint_bwd_a") g_hint_bwd_a = value;
  else if (n == "hint_bwd_b") g_hint_bwd_b = value;
  else return O3V_ERR_INVALID_ARG;
  return O3V_OK;
}

extern "C" size_t o3v_lmhead_fwd_workspace_bytes(int64_t T, int64_t V, int64_t H) {
  if (T <= 0 || V <= 0) return 0;
  // one partial triple per token and per vocab group of the plan o3v_lmhead_fwd will make for this shape under the
  // current tunables (4-16 groups; the worst case, one group per 256-column tile, was 0.9 GB at T = 131072)
  GemmParams p = {};
  p.M = T; p.N = V; p.K = H;
  plan_tiles(p, g_cta_fwd, fwd_groups(T, V, g_cta_fwd));
  return (size_t)((int64_t)p.num_n_groups * 3 * T) * sizeof(float);
}

extern "C" int o3v_lmhead_fwd(const void* hidden, const void* weight, const int64_t* targets,
                              int64_t T, int64_t V, int64_t H, int64_t v_offset,
                              float* stats, void* logits, int64_t ld_logits,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (!hidden || !weight || !targets || !stats || !workspace) return O3V_ERR_INVALID_ARG;
  if (T <= 0 || V <= 0 || H <= 0 || v_offset < 0) return O3V_ERR_INVALID_ARG;
  if (H % 64 != 0) return O3V_ERR_SHAPE;
  if (T > 0x7fffffffLL - 512 || V > 0x7fffffffLL - 512) return O3V_ERR_SHAPE;
  if (logits) {
    if (V % 8 != 0 || ld_logits % 8 != 0 || ld_logits < V) return O3V_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(logits) & 15u) return O3V_ERR_ALIGNMENT;
  }
  int rc = check_device();
  if (rc) return rc;
  const int ncta = g_cta_fwd;
  GemmParams p = {};
  p.M = T; p.N = V; p.K = H;
  plan_tiles(p, ncta, fwd_groups(T, V, ncta));
  if (workspace_bytes < (size_t)p.num_n_groups * 3 * T * sizeof(float)) return O3V_ERR_WORKSPACE;
  p.targets = targets; p.v_offset = v_offset; p.parts = reinterpret_cast<float*>(workspace);
  p.logits = reinterpret_cast<__nv_bfloat16*>(logits); p.ld_logits = ld_logits;
  p.hint_a = g_hint_fwd_a; p.hint_b = g_hint_fwd_b; p.hint_store = g_hint_fwd_store;
  CUtensorMap tmA, tmB, tmC = {};
  if ((rc = make_tmap_bf16(&tmA, hidden, H, T, H, 128))) return rc;
  if ((rc = make_tmap_bf16(&tmB, weight, H, V, H, 256 / ncta))) return rc;
  if (logits && (rc = make_tmap_bf16(&tmC, logits, V, T, ld_logits, 32))) return rc;   // store box: 32 rows x 64 cols
  cudaStream_t st = (cudaStream_t)stream;
  rc = (ncta == 1) ? launch_gemm<false, false, 1, EPI_STATS>(tmA, tmB, tmC, p, st)
                   : launch_gemm<false, false, 2, EPI_STATS>(tmA, tmB, tmC, p, st);
  if (rc) return rc;
  merge_stats_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, st>>>(p.parts, p.num_n_groups, T, stats, nullptr, nullptr);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_merge_stats(const float* parts, int64_t P, int64_t T, float* logp, float* lse,
                                      void* stream) {
  if (!parts || (!logp && !lse) || P <= 0 || T <= 0) return O3V_ERR_INVALID_ARG;
  int rc = check_device();
  if (rc) return rc;
  merge_stats_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, (cudaStream_t)stream>>>(parts, P, T, nullptr, logp, lse);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_merge_stats_peers(const float* const* part_ptrs, int64_t P, int64_t row_stride, int64_t T,
                                            float* logp, float* lse, void* stream) {
  if (!part_ptrs || (!logp && !lse) || P <= 0 || P > 16 || T <= 0 || row_stride < T) return O3V_ERR_INVALID_ARG;
  int rc = check_device();
  if (rc) return rc;
  PeerPtrs ptrs = {};
  for (int64_t i = 0; i < P; ++i) {
    if (!part_ptrs[i]) return O3V_ERR_INVALID_ARG;
    ptrs.p[i] = part_ptrs[i];
  }
  merge_stats_peers_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, (cudaStream_t)stream>>>(ptrs, (int)P, row_stride, T, logp, lse);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_allreduce_bf16_peers(void* const* bufs, int64_t P, int64_t rank, int64_t n_elems, int32_t num_ctas,
                                        void* stream) {
  if (!bufs || P <= 0 || P > 16 || rank < 0 || rank >= P || n_elems < 0 || (n_elems % 8) != 0) return O3V_ERR_INVALID_ARG;
  int rc = check_device();
  if (rc) return rc;
  if (n_elems == 0 || P == 1) return O3V_OK;
  PeerBufs pb = {};
  for (int64_t i = 0; i < P; ++i) {
    if (!bufs[i] || (reinterpret_cast<uintptr_t>(bufs[i]) & 15u)) return O3V_ERR_ALIGNMENT;
    pb.p[i] = reinterpret_cast<uint4*>(bufs[i]);
  }
  const int64_t nvec = n_elems / 8;
  const int64_t v0 = rank * nvec / P, v1 = (rank + 1) * nvec / P;
  if (v1 <= v0) return O3V_OK;
  int ctas = num_ctas > 0 ? num_ctas : 2 * num_sms();
  ctas = (int)std::min<int64_t>(ctas, ceil_div(v1 - v0, kArThreads));
  allreduce_bf16_peers_kernel<<<ctas, kArThreads, 0, (cudaStream_t)stream>>>(pb, (int)P, v0, v1);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_dlogits(void* logits, int64_t T, int64_t V, int64_t ld_logits, const float* lse,
                                  const float* grad_logp, const int64_t* targets, int64_t v_offset,
                                  void* stream) {
  if (!logits || !lse || !grad_logp || !targets || T <= 0 || V <= 0) return O3V_ERR_INVALID_ARG;
  if (V % 8 != 0 || ld_logits % 8 != 0 || ld_logits < V) return O3V_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(logits) & 15u) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  // several waves of one-row CTAs; rows are 300 KB each so a CTA streams plenty of bytes
  const int rows_per_cta = 1;
  dlogits_kernel<<<(unsigned)ceil_div(T, rows_per_cta), kDlThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(logits), T, V, ld_logits, lse, grad_logp, targets, v_offset, rows_per_cta);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

// softmax-backward parameters of the fused variants (nullptr = plain GEMM on a ready-made P)
struct XformArgs { const void* rows; };

static int bwd_dhidden_impl(const void* dlogits, int64_t ld_dlogits, const void* weight, int64_t T, int64_t V, int64_t H,
                            void* d_hidden, int32_t out_is_fp32, const XformArgs* x, void* stream) {
  if (!dlogits || !weight || !d_hidden || T <= 0 || V <= 0 || H <= 0) return O3V_ERR_INVALID_ARG;
  if (H % 8 != 0 || ld_dlogits % 8 != 0 || ld_dlogits < V) return O3V_ERR_SHAPE;
  if (T > 0x7fffffffLL - 512 || V > 0x7fffffffLL - 512) return O3V_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(d_hidden) & 15u) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  const int ncta = g_cta_bwd;
  GemmParams p = {};
  p.M = T; p.N = H; p.K = V;                       // dH[t,h] = sum_v P[t,v] W[v,h]
  const bool wide = (ncta == 2 && g_bwd_wide);
  plan_tiles(p, ncta, 1 << 20, wide ? 2 : 1);       // one n-tile per item
  p.out = d_hidden; p.ld_out = H; p.out_fp32 = out_is_fp32 ? 1 : 0; p.m_fast = g_dh_mfast;
  p.prefetch_a = g_bwd_prefetch;
  p.hint_a = g_hint_bwd_a; p.hint_b = g_hint_bwd_b;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16(&tmA, dlogits, V, T, ld_dlogits, 128))) return rc;   // A = P, K-major (K = V)
  if ((rc = make_tmap_bf16(&tmB, weight, H, V, H, 64))) return rc;              // B = W, MN-major (N = H contiguous)
  cudaStream_t st = (cudaStream_t)stream;
  if (g_bwd_sync && (rc = acquire_wave_sync(&p.wave_sync, st))) return rc;
  if (x) {
    if (!wide) return O3V_ERR_UNSUPPORTED_MODE;        // the transform runs on the idle epilogue warps of 256x512 pair tiles
    p.x_rows = reinterpret_cast<const int4*>(x->rows); p.x_tokens = T;
    return launch_gemm<false, true, 2, EPI_STORE, 2, true>(tmA, tmB, tmA, p, st);
  }
  if (wide) return launch_gemm<false, true, 2, EPI_STORE, 2>(tmA, tmB, tmA, p, st);
  return (ncta == 1) ? launch_gemm<false, true, 1, EPI_STORE>(tmA, tmB, tmA, p, st)
                     : launch_gemm<false, true, 2, EPI_STORE>(tmA, tmB, tmA, p, st);
}

extern "C" int o3v_lmhead_bwd_dhidden(const void* dlogits, int64_t ld_dlogits, const void* weight,
                                      int64_t T, int64_t V, int64_t H, void* d_hidden, int32_t out_is_fp32,
                                      void* stream) {
  return bwd_dhidden_impl(dlogits, ld_dlogits, weight, T, V, H, d_hidden, out_is_fp32, nullptr, stream);
}

__global__ void softmax_bwd_rows_kernel(const float* __restrict__ lse, const float* __restrict__ g,
                                        const int64_t* __restrict__ targets, int64_t v_offset, int64_t V, int64_t T,
                                        SoftmaxBwdRow* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  SoftmaxBwdRow r;
  r.g = g[t];
  r.c = r.g != 0.f ? __log2f(fabsf(r.g)) - lse[t] * kLog2e : 0.f;
  const int64_t col = targets[t] - v_offset;
  r.tcol = (col >= 0 && col < V) ? (int32_t)col : -1;
  r.pad = 0;
  out[t] = r;
}

extern "C" int o3v_lmhead_softmax_bwd_rows(const float* lse, const float* grad_logp, const int64_t* targets,
                                           int64_t v_offset, int64_t V, int64_t T, void* rows, void* stream) {
  if (!lse || !grad_logp || !targets || !rows || v_offset < 0 || V <= 0 || T <= 0) return O3V_ERR_INVALID_ARG;
  if (V > 0x7fffffffLL - 512) return O3V_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(rows) & 15u) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  static_assert(sizeof(SoftmaxBwdRow) == 16, "record layout");
  softmax_bwd_rows_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, (cudaStream_t)stream>>>(
      lse, grad_logp, targets, v_offset, V, T, reinterpret_cast<SoftmaxBwdRow*>(rows));
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_bwd_dhidden_fused(const void* logits, int64_t ld_logits, const void* rows,
                                            const void* weight, int64_t T, int64_t V, int64_t H, void* d_hidden,
                                            int32_t out_is_fp32, void* stream) {
  if (!rows || (reinterpret_cast<uintptr_t>(rows) & 15u)) return O3V_ERR_INVALID_ARG;
  XformArgs x = {rows};
  return bwd_dhidden_impl(logits, ld_logits, weight, T, V, H, d_hidden, out_is_fp32, &x, stream);
}

static int bwd_dweight_impl(const void* dlogits, int64_t ld_dlogits, const void* hidden, int64_t T, int64_t V, int64_t H,
                            float* d_weight, int32_t accumulate, const XformArgs* x, void* stream) {
  if (!dlogits || !hidden || !d_weight || T <= 0 || V <= 0 || H <= 0) return O3V_ERR_INVALID_ARG;
  if (H % 8 != 0 || ld_dlogits % 8 != 0 || ld_dlogits < V) return O3V_ERR_SHAPE;
  if (T > 0x7fffffffLL - 512 || V > 0x7fffffffLL - 512) return O3V_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(d_weight) & 15u) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  const int ncta = g_cta_bwd;
  GemmParams p = {};
  p.M = V; p.N = H; p.K = T;                       // dW[v,h] = sum_t P[t,v] hidden[t,h]
  const bool wide = (ncta == 2 && g_bwd_wide);
  plan_tiles(p, ncta, 1 << 20, wide ? 2 : 1);
  p.out = d_weight; p.ld_out = H; p.out_fp32 = 1; p.accumulate = accumulate ? 1 : 0; p.m_fast = g_dw_mfast;
  p.prefetch_a = g_bwd_prefetch;
  p.hint_a = g_hint_bwd_a; p.hint_b = g_hint_bwd_b;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16(&tmA, dlogits, V, T, ld_dlogits, 64))) return rc;    // A = P^T, MN-major (M = V contiguous)
  if ((rc = make_tmap_bf16(&tmB, hidden, H, T, H, 64))) return rc;              // B = hidden, MN-major
  cudaStream_t st = (cudaStream_t)stream;
  if (g_bwd_sync && (rc = acquire_wave_sync(&p.wave_sync, st))) return rc;
  if (x) {
    if (!wide) return O3V_ERR_UNSUPPORTED_MODE;
    p.x_rows = reinterpret_cast<const int4*>(x->rows); p.x_tokens = T;
    return launch_gemm<true, true, 2, EPI_ACCUM, 2, true>(tmA, tmB, tmA, p, st);
  }
  if (wide) return launch_gemm<true, true, 2, EPI_ACCUM, 2>(tmA, tmB, tmA, p, st);
  return (ncta == 1) ? launch_gemm<true, true, 1, EPI_ACCUM>(tmA, tmB, tmA, p, st)
                     : launch_gemm<true, true, 2, EPI_ACCUM>(tmA, tmB, tmA, p, st);
}

extern "C" int o3v_lmhead_bwd_dweight(const void* dlogits, int64_t ld_dlogits, const void* hidden,
                                      int64_t T, int64_t V, int64_t H, float* d_weight, int32_t accumulate,
                                      void* stream) {
  return bwd_dweight_impl(dlogits, ld_dlogits, hidden, T, V, H, d_weight, accumulate, nullptr, stream);
}

extern "C" int o3v_lmhead_bwd_dweight_fused(const void* logits, int64_t ld_logits, const void* rows,
                                            const void* hidden, int64_t T, int64_t V, int64_t H, float* d_weight,
                                            int32_t accumulate, void* stream) {
  if (!rows || (reinterpret_cast<uintptr_t>(rows) & 15u)) return O3V_ERR_INVALID_ARG;
  XformArgs x = {rows};
  return bwd_dweight_impl(logits, ld_logits, hidden, T, V, H, d_weight, accumulate, &x, stream);
}

extern "C" int o3v_debug_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N,
                              int64_t K, int32_t a_mn, int32_t b_mn, void* out, int64_t ld_out,
                              int32_t out_fp32, int32_t accumulate, void* stream) {
  if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return O3V_ERR_INVALID_ARG;
  int rc = check_device();
  if (rc) return rc;
  const int ncta = g_cta_bwd;
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K;
  plan_tiles(p, ncta, (int)ceil_div(N, 256));
  p.out = out; p.ld_out = ld_out; p.out_fp32 = (out_fp32 || accumulate) ? 1 : 0; p.accumulate = accumulate ? 1 : 0;
  CUtensorMap tmA, tmB;
  if (a_mn) { if ((rc = make_tmap_bf16(&tmA, A, M, K, lda, 64))) return rc; }
  else      { if ((rc = make_tmap_bf16(&tmA, A, K, M, lda, 128))) return rc; }
  if (b_mn) { if ((rc = make_tmap_bf16(&tmB, B, N, K, ldb, 64))) return rc; }
  else      { if ((rc = make_tmap_bf16(&tmB, B, K, N, ldb, 256 / ncta))) return rc; }
  cudaStream_t st = (cudaStream_t)stream;
#define O3V_DISPATCH(AM, BM_, EPI)                                                    \
  return (ncta == 1) ? launch_gemm<AM, BM_, 1, EPI>(tmA, tmB, tmA, p, st)             \
                     : launch_gemm<AM, BM_, 2, EPI>(tmA, tmB, tmA, p, st)
  if (accumulate) {
    if (a_mn && b_mn) { O3V_DISPATCH(true, true, EPI_ACCUM); }
    if (!a_mn && !b_mn) { O3V_DISPATCH(false, false, EPI_ACCUM); }
    return O3V_ERR_INVALID_ARG;
  }
  if (!a_mn && b_mn) { O3V_DISPATCH(false, true, EPI_STORE); }
  if (!a_mn && !b_mn) { O3V_DISPATCH(false, false, EPI_STORE); }
  return O3V_ERR_INVALID_ARG;
#undef O3V_DISPATCH
}
