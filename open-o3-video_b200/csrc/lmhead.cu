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
static int g_fwd_rotate = 1;    // K1: rotate the vocab group a worker takes from wave to wave
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
__device__ unsigned int g_wave_sync[kSyncSlots][2];     // {arrivals, abandoned}

static int acquire_wave_sync(unsigned int** out, cudaStream_t st) {
  static std::atomic<unsigned int> next{0};
  unsigned int* base = nullptr;
  O3V_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_wave_sync));
  unsigned int* slot = base + 2 * (next.fetch_add(1) % kSyncSlots);
  O3V_CUDA_TRY(cudaMemsetAsync(slot, 0, 2 * sizeof(unsigned int), st));
  *out = slot;
  return O3V_OK;
}

// Test helper: `num_ctas` CTAs that each hold `smem_bytes` of shared memory and spin for `clocks` SM clocks, to occupy
// SMs beside a persistent GEMM on another stream (tests/test_gpu_configs.py: the wave rendezvous must degrade, not trap).
__global__ void occupy_sms_kernel(long long clocks) {
  extern __shared__ uint8_t occupy_smem[];
  occupy_smem[threadIdx.x] = (uint8_t)threadIdx.x;
  const long long t0 = clock64();
  while (clock64() - t0 < clocks) __nanosleep(1000);
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
  p.num_n_groups = groups;
  p.tiles_per_group = (int32_t)ceil_div(p.num_n_tiles, groups);
}

// Number of vocab splits per m-block in K1.  Few splits keep the concurrently swept A panels
// (hidden rows) L2-resident and W re-reads low; enough splits fill the persistent grid.  Four is the measured
// optimum when the work fills its waves (profiles/r1_sweep_T32768_7b.log); another count (up to 16) is taken only
// when the wave arithmetic says it saves >= 4 %: time ~ ceil(m_blocks * g / workers) waves x ceil(tiles / g) tiles,
// e.g. 272 m-blocks x 594 tiles: g = 4 -> 8 x 149, g = 7 -> 13 x 85 (-7 %).
static int fwd_groups(int64_t T, int64_t V, int ncta) {
  if (g_fwd_groups > 0) return g_fwd_groups;
  const int64_t num_m = ceil_div(T, 128 * ncta);
  const int64_t n_tiles = ceil_div(V, 256);
  const int64_t workers = num_sms() / ncta;
  int64_t g = 4;
  while (num_m * g < 2 * workers && g < n_tiles) g *= 2;     // small T: split the vocab further
  if (g > n_tiles) g = n_tiles;
  if (g == 4 && num_m * g >= 4 * workers) {
    auto cost = [&](int64_t gg) { return ceil_div(num_m * gg, workers) * ceil_div(n_tiles, gg); };
    int64_t best = g, best_cost = cost(g);
    for (int64_t gg = 5; gg <= 16 && gg <= n_tiles; ++gg)
      if (cost(gg) * 100 < best_cost * 99) { best = gg; best_cost = cost(gg); }   // near-ties: the smaller count
    if (best_cost * 100 <= cost(g) * 96) g = best;
  }
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

// Reduce-scatter by PULL: out[i] = sum_p bufs[p][v0 + i] for this rank's vectors, fp32 in rank order, one bf16 rounding,
// stored LOCALLY only (no write fan-out: half the NVLink bytes of the all-reduce above).  Same footprint as the
// all-reduce kernel (no shared memory, few registers) so that it runs beside the persistent dW GEMM.
__global__ void __launch_bounds__(kArThreads, 4)
reduce_scatter_bf16_peers_kernel(const PeerBufs bufs, int P, int64_t v0, int64_t n_vecs, uint4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kArThreads;
  for (int64_t i = (int64_t)blockIdx.x * kArThreads + threadIdx.x; i < n_vecs; i += stride) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p0 = 0; p0 < P; p0 += 4) {
      uint4 a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + j < P) a[j] = __ldcv(bufs.p[p0 + j] + v0 + i);   // four NVLink loads in flight
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + j < P) {                                        // fixed rank order: deterministic
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
    out[i] = r;
  }
}

// Sum of the P slot copies [P][slot_rows][H] (bf16, this rank's buffer: slot p was written by rank p's K2a epilogue over
// NVLink) into out[rows][H]: fp32 in rank order (deterministic), one bf16 rounding.  Local HBM traffic only.
__global__ void __launch_bounds__(256)
sum_slots_bf16_kernel(const uint4* __restrict__ slots, int P, int64_t slot_vecs, int64_t n_vecs, uint4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vecs; i += stride) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < P; ++p) {
      const uint4 a = __ldcs(slots + (int64_t)p * slot_vecs + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        s[2 * k] += f.x; s[2 * k + 1] += f.y;
      }
    }
    uint4 r;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = __floats2bfloat162_rn(s[2 * k], s[2 * k + 1]);
    out[i] = r;
  }
}

// out[i] += sum_s slabs[s][i] in slab order (the K-split partial sums of the dW tail tiles: deterministic)
__global__ void __launch_bounds__(256)
add_slabs_f32_kernel(const float4* __restrict__ slabs, int S, int64_t n_vecs, float4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vecs; i += stride) {
    float4 acc = out[i];
    for (int sidx = 0; sidx < S; ++sidx) {
      const float4 v = __ldcs(slabs + (int64_t)sidx * n_vecs + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[i] = acc;
  }
}

// ------------------------------------------------------------------------------------
// exp-store backward: P = a * E + g * onehot, so  dH = a * (E . W) + g * W[tcol]   (K2a epilogue)
//                                                 dW = E^T . (a * hidden) + scatter_t g_t * hidden_t into row tcol_t
// ------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ lse, const float* __restrict__ g,
                                    const int64_t* __restrict__ targets, const float* __restrict__ row_ref,
                                    int64_t v_offset, int64_t V, int64_t T, SoftmaxEpiRow* __restrict__ out,
                                    int64_t* __restrict__ sort_key) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  SoftmaxEpiRow r;
  r.g = g[t];
  r.a = r.g != 0.f ? -r.g * exp2f((row_ref[t] - lse[t]) * kLog2e) : 0.f;
  const int64_t col = targets[t] - v_offset;
  r.tcol = (col >= 0 && col < V) ? (int32_t)col : -1;
  r.pad = 0;
  out[t] = r;
  // key of the one-hot scatter: tokens that contribute to a row of dW sort by that row, the rest go last
  if (sort_key) sort_key[t] = (r.g != 0.f && r.tcol >= 0) ? (int64_t)r.tcol : (int64_t)0x7fffffffffffLL;
}

// out[t, :] = a_t * hidden[t, :]  (bf16, 16 bytes per thread)
__global__ void __launch_bounds__(256)
scale_rows_kernel(const uint4* __restrict__ hidden, const SoftmaxEpiRow* __restrict__ rows, int64_t T, int64_t vec_per_row,
                  uint4* __restrict__ out) {
  const int64_t n = T * vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = rows[i / vec_per_row].a;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (a != 0.f) {
      x = hidden[i];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        h[k] = __floats2bfloat162_rn(a * f.x, a * f.y);
      }
    }
    out[i] = x;
  }
}

// One CTA per position i of the sorted token order; the CTA at the first token of a run of equal target columns sums
// g_t * hidden[t, :] over the run sequentially in fp32 (deterministic) and adds it to dW[tcol, :] (single writer per row).
__global__ void __launch_bounds__(128)
onehot_scatter_kernel(const int64_t* __restrict__ order, const SoftmaxEpiRow* __restrict__ rows,
                      const __nv_bfloat16* __restrict__ hidden, int64_t T, int64_t H, float* __restrict__ dW) {
  const int64_t i = blockIdx.x;
  const int64_t t0 = order[i];
  const SoftmaxEpiRow r0 = rows[t0];
  if (r0.g == 0.f || r0.tcol < 0) return;
  if (i > 0) {
    const SoftmaxEpiRow rp = rows[order[i - 1]];
    if (rp.g != 0.f && rp.tcol == r0.tcol) return;          // not the first token of its run
  }
  float* dst = dW + (int64_t)r0.tcol * H;
  for (int64_t h0 = (int64_t)threadIdx.x * 8; h0 < H; h0 += 128 * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t j = i; j < T; ++j) {
      const int64_t t = order[j];
      const SoftmaxEpiRow r = rows[t];
      if (r.g == 0.f || r.tcol != r0.tcol) break;
      const uint4 x = *reinterpret_cast<const uint4*>(hidden + t * H + h0);
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(hh[k]);
        acc[2 * k] = fmaf(r.g, f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(r.g, f.y, acc[2 * k + 1]);
      }
    }
    float4* d4 = reinterpret_cast<float4*>(dst + h0);
    float4 lo = d4[0], hi = d4[1];
    lo.x += acc[0]; lo.y += acc[1]; lo.z += acc[2]; lo.w += acc[3];
    hi.x += acc[4]; hi.y += acc[5]; hi.z += acc[6]; hi.w += acc[7];
    d4[0] = lo; d4[1] = hi;
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
  else if (n == "fwd_rotate") g_fwd_rotate = value ? 1 : 0;
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

static int lmhead_fwd_impl(const void* hidden, const void* weight, const int64_t* targets,
                           int64_t T, int64_t V, int64_t H, int64_t v_offset,
                           float* stats, void* logits, int64_t ld_logits, const float* row_ref, const int32_t* row_keep,
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
  p.row_ref = row_ref; p.row_keep = row_keep;
  {
    // group rotation per wave of the persistent grid (only when a wave holds whole m-blocks)
    int sms = num_sms();
    if (g_max_ctas > 0 && g_max_ctas < sms) sms = g_max_ctas;
    const int64_t workers = std::min<int64_t>(sms / ncta, (int64_t)p.num_m_blocks * p.num_n_groups);
    if (g_fwd_rotate && p.num_n_groups > 1 && workers >= p.num_n_groups && workers % p.num_n_groups == 0)
      p.rotate_groups = (int32_t)(workers / p.num_n_groups);
  }
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

extern "C" int o3v_lmhead_fwd(const void* hidden, const void* weight, const int64_t* targets,
                              int64_t T, int64_t V, int64_t H, int64_t v_offset,
                              float* stats, void* logits, int64_t ld_logits,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return lmhead_fwd_impl(hidden, weight, targets, T, V, H, v_offset, stats, logits, ld_logits, nullptr, nullptr, workspace,
                         workspace_bytes, stream);
}

extern "C" int o3v_lmhead_fwd_exp(const void* hidden, const void* weight, const int64_t* targets,
                                  int64_t T, int64_t V, int64_t H, int64_t v_offset,
                                  float* stats, void* expz, int64_t ld_expz, const float* row_ref, const int32_t* row_keep,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  if (!expz || !row_ref) return O3V_ERR_INVALID_ARG;
  return lmhead_fwd_impl(hidden, weight, targets, T, V, H, v_offset, stats, expz, ld_expz, row_ref, row_keep, workspace,
                         workspace_bytes, stream);
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

extern "C" int o3v_reduce_scatter_bf16_peers(void* const* bufs, int64_t P, int64_t elem_offset, int64_t n_elems,
                                             void* out, int32_t num_ctas, void* stream) {
  if (!bufs || !out || P <= 0 || P > 16 || elem_offset < 0 || n_elems < 0 || (n_elems % 8) != 0 || (elem_offset % 8) != 0)
    return O3V_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(out) & 15u) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  if (n_elems == 0) return O3V_OK;
  PeerBufs pb = {};
  for (int64_t i = 0; i < P; ++i) {
    if (!bufs[i] || (reinterpret_cast<uintptr_t>(bufs[i]) & 15u)) return O3V_ERR_ALIGNMENT;
    pb.p[i] = reinterpret_cast<uint4*>(bufs[i]);
  }
  const int64_t nvec = n_elems / 8;
  int ctas = num_ctas > 0 ? num_ctas : 2 * num_sms();
  ctas = (int)std::min<int64_t>(ctas, ceil_div(nvec, kArThreads));
  reduce_scatter_bf16_peers_kernel<<<ctas, kArThreads, 0, (cudaStream_t)stream>>>(pb, (int)P, elem_offset / 8, nvec,
                                                                                  reinterpret_cast<uint4*>(out));
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
// reduce-scatter destination of K2a (nullptr = local store)
struct ScatterArgs { void* const* slot_ptrs; int64_t P, rank, rows_per_owner, slot_rows, row0; };
// exp-store path: the A operand is E and the softmax backward happens in the K2a epilogue (nullptr = A is P)
struct SoftmaxEpiArgs { const void* rows; };

static int bwd_dhidden_impl(const void* dlogits, int64_t ld_dlogits, const void* weight, int64_t T, int64_t V, int64_t H,
                            void* d_hidden, int32_t out_is_fp32, const XformArgs* x, void* stream,
                            const ScatterArgs* sc = nullptr, const SoftmaxEpiArgs* sm = nullptr) {
  if (!dlogits || !weight || (!d_hidden && !sc) || T <= 0 || V <= 0 || H <= 0) return O3V_ERR_INVALID_ARG;
  if (H % 8 != 0 || ld_dlogits % 8 != 0 || ld_dlogits < V) return O3V_ERR_SHAPE;
  if (T > 0x7fffffffLL - 512 || V > 0x7fffffffLL - 512) return O3V_ERR_SHAPE;
  if (d_hidden && (reinterpret_cast<uintptr_t>(d_hidden) & 15u)) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  const int ncta = g_cta_bwd;
  GemmParams p = {};
  p.M = T; p.N = H; p.K = V;                       // dH[t,h] = sum_v P[t,v] W[v,h]
  const bool wide = (ncta == 2 && g_bwd_wide);
  plan_tiles(p, ncta, 1 << 20, wide ? 2 : 1);       // one n-tile per item
  p.out = d_hidden; p.ld_out = H; p.out_fp32 = out_is_fp32 ? 1 : 0; p.m_fast = g_dh_mfast;
  p.prefetch_a = g_bwd_prefetch;
  if (sm) {
    if (!sm->rows || (reinterpret_cast<uintptr_t>(sm->rows) & 15u) || x) return O3V_ERR_INVALID_ARG;
    p.sm_rows = reinterpret_cast<const int4*>(sm->rows);
    p.sm_weight = reinterpret_cast<const __nv_bfloat16*>(weight);
  }
  if (sc) {
    if (out_is_fp32 || sc->P < 1 || sc->P > 16 || sc->rank < 0 || sc->rank >= sc->P || sc->rows_per_owner < 1 ||
        sc->slot_rows < sc->rows_per_owner || sc->row0 < 0 || !sc->slot_ptrs)
      return O3V_ERR_INVALID_ARG;
    if (sc->row0 + T > sc->P * sc->rows_per_owner) return O3V_ERR_SHAPE;      // rows beyond the last owner's range
    for (int64_t i = 0; i < sc->P; ++i) {
      if (!sc->slot_ptrs[i] || (reinterpret_cast<uintptr_t>(sc->slot_ptrs[i]) & 15u)) return O3V_ERR_ALIGNMENT;
      p.peer_out[i] = reinterpret_cast<__nv_bfloat16*>(sc->slot_ptrs[i]);
    }
    p.n_peers = (int32_t)sc->P; p.my_rank = (int32_t)sc->rank;
    p.rows_per_owner = sc->rows_per_owner; p.slot_rows = sc->slot_rows; p.row0 = sc->row0;
  }
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

extern "C" int o3v_lmhead_bwd_dhidden_scatter(const void* dlogits, int64_t ld_dlogits, const void* weight,
                                              int64_t T, int64_t V, int64_t H, void* const* slot_ptrs, int64_t P,
                                              int64_t rank, int64_t rows_per_owner, int64_t slot_rows, int64_t row0,
                                              void* stream) {
  ScatterArgs sc = {slot_ptrs, P, rank, rows_per_owner, slot_rows, row0};
  return bwd_dhidden_impl(dlogits, ld_dlogits, weight, T, V, H, nullptr, 0, nullptr, stream, &sc);
}

extern "C" int o3v_lmhead_softmax_rows(const float* lse, const float* grad_logp, const int64_t* targets,
                                       const float* row_ref, int64_t v_offset, int64_t V, int64_t T, void* rows,
                                       int64_t* sort_key, void* stream) {
  if (!lse || !grad_logp || !targets || !row_ref || !rows || v_offset < 0 || V <= 0 || T <= 0) return O3V_ERR_INVALID_ARG;
  if (V > 0x7fffffffLL - 512) return O3V_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(rows) & 15u) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  static_assert(sizeof(SoftmaxEpiRow) == 16, "record layout");
  softmax_rows_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, (cudaStream_t)stream>>>(
      lse, grad_logp, targets, row_ref, v_offset, V, T, reinterpret_cast<SoftmaxEpiRow*>(rows), sort_key);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_bwd_dhidden_exp(const void* expz, int64_t ld_expz, const void* rows, const void* weight,
                                          int64_t T, int64_t V, int64_t H, void* d_hidden, int32_t out_is_fp32,
                                          void* const* slot_ptrs, int64_t P, int64_t rank, int64_t rows_per_owner,
                                          int64_t slot_rows, int64_t row0, void* stream) {
  SoftmaxEpiArgs sm = {rows};
  if (slot_ptrs) {
    ScatterArgs sc = {slot_ptrs, P, rank, rows_per_owner, slot_rows, row0};
    return bwd_dhidden_impl(expz, ld_expz, weight, T, V, H, nullptr, 0, nullptr, stream, &sc, &sm);
  }
  return bwd_dhidden_impl(expz, ld_expz, weight, T, V, H, d_hidden, out_is_fp32, nullptr, stream, nullptr, &sm);
}

extern "C" int o3v_sum_slots_bf16(const void* slots, int64_t P, int64_t slot_rows, int64_t rows, int64_t H, void* out,
                                  int32_t num_ctas, void* stream) {
  if (!slots || !out || P < 1 || P > 16 || rows < 0 || slot_rows < rows || H <= 0 || (H % 8) != 0) return O3V_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(slots) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  if (rows == 0) return O3V_OK;
  const int64_t n_vecs = rows * H / 8, slot_vecs = slot_rows * H / 8;
  int ctas = num_ctas > 0 ? num_ctas : 2 * num_sms();
  ctas = (int)std::min<int64_t>(ctas, ceil_div(n_vecs, 256));
  sum_slots_bf16_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(slots), (int)P, slot_vecs,
                                                                n_vecs, reinterpret_cast<uint4*>(out));
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

extern "C" int o3v_add_slabs_f32(const float* slabs, int64_t S, int64_t n_elems, float* out, void* stream) {
  if (!slabs || !out || S < 0 || n_elems < 0 || (n_elems % 4) != 0) return O3V_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(slabs) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  if (S == 0 || n_elems == 0) return O3V_OK;
  const int64_t nvec = n_elems / 4;
  const int ctas = (int)std::min<int64_t>(4 * (int64_t)num_sms(), ceil_div(nvec, 256));
  add_slabs_f32_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(slabs), (int)S, nvec,
                                                               reinterpret_cast<float4*>(out));
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_bwd_dweight_exp(const void* expz, int64_t ld_expz, const void* rows, const int64_t* order,
                                          const void* hidden, int64_t T, int64_t V, int64_t H, float* d_weight,
                                          int32_t accumulate, void* scaled_hidden, void* stream) {
  if (!rows || !order || !scaled_hidden || !hidden || T <= 0 || H <= 0 || (H % 8) != 0) return O3V_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(rows) & 15u) || (reinterpret_cast<uintptr_t>(scaled_hidden) & 15u) ||
      (reinterpret_cast<uintptr_t>(hidden) & 15u))
    return O3V_ERR_ALIGNMENT;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nvec = T * (H / 8);
  const int ctas = (int)std::min<int64_t>(8 * (int64_t)num_sms(), ceil_div(nvec, 256));
  scale_rows_kernel<<<ctas, 256, 0, st>>>(reinterpret_cast<const uint4*>(hidden),
                                          reinterpret_cast<const SoftmaxEpiRow*>(rows), T, H / 8,
                                          reinterpret_cast<uint4*>(scaled_hidden));
  O3V_LAUNCH_CHECK();
  rc = bwd_dweight_impl(expz, ld_expz, scaled_hidden, T, V, H, d_weight, accumulate, nullptr, stream);
  if (rc) return rc;
  onehot_scatter_kernel<<<(unsigned)T, 128, 0, st>>>(order, reinterpret_cast<const SoftmaxEpiRow*>(rows),
                                                     reinterpret_cast<const __nv_bfloat16*>(hidden), T, H, d_weight);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" int o3v_lmhead_bwd_dweight_fused(const void* logits, int64_t ld_logits, const void* rows,
                                            const void* hidden, int64_t T, int64_t V, int64_t H, float* d_weight,
                                            int32_t accumulate, void* stream) {
  if (!rows || (reinterpret_cast<uintptr_t>(rows) & 15u)) return O3V_ERR_INVALID_ARG;
  XformArgs x = {rows};
  return bwd_dweight_impl(logits, ld_logits, hidden, T, V, H, d_weight, accumulate, &x, stream);
}

extern "C" int o3v_debug_occupy_sms(int32_t num_ctas, int32_t smem_bytes, int64_t clocks, void* stream) {
  if (num_ctas < 1 || smem_bytes < 0 || smem_bytes > 227 * 1024 || clocks < 0) return O3V_ERR_INVALID_ARG;
  int rc = check_device();
  if (rc) return rc;
  O3V_CUDA_TRY(cudaFuncSetAttribute(occupy_sms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  occupy_sms_kernel<<<num_ctas, 64, smem_bytes, (cudaStream_t)stream>>>((long long)clocks);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
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
