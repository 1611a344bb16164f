// Warp-specialised, persistent tcgen05 GEMM for the lm_head hot path (sm_100a).
//
//   D[M, N] = A[M, K] . B[N, K]^T      bf16 x bf16 -> fp32 accumulators in TMEM
//
// One kernel template serves the three contractions of the path; they differ in operand
// majorness (chosen so that NO tensor is ever transposed in HBM) and in the epilogue:
//
//   K1  logits tile = hidden . W^T        A K-major,  B K-major   EPI_STATS
//       online (max, sum-exp) + target-logit capture per token row, carried in registers
//       across the n-tiles of a work item; optional bf16 store of the tile for the backward
//   K2a dHidden = P . W                   A K-major,  B MN-major  EPI_STORE (bf16 / fp32)
//   K2b dW (+)= P^T . hidden              A MN-major, B MN-major  EPI_ACCUM (fp32 RMW)
//
// Structure (256 threads, 1 CTA per SM, optionally a cta_group::2 pair per tile):
//   warp 0 / lane 0   TMA producer: cp.async.bulk.tensor 2-D boxes, 128B swizzle, into a
//                     kStages-deep smem ring guarded by full/empty mbarriers
//   warp 1 / lane 0   MMA issuer: tcgen05.mma (M=128 or 256 with cta_group::2, N=256, K=16),
//                     tcgen05.commit releases smem stages and publishes accumulators
//   warp 2            TMEM allocator (512 columns = 2 accumulator stages of 256)
//   warps 4..7        epilogue: tcgen05.ld 32x32b (one accumulator row per thread), overlaps
//                     the next tile's MMAs through the double-buffered accumulator
// Work items are statically strided over the persistent grid; an item is an (m-block,
// n-group) pair, n-groups fastest so that concurrently running CTAs share A panels in L2.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace o3v {

enum : int { EPI_STATS = 0, EPI_STORE = 1, EPI_ACCUM = 2 };

struct GemmParams {
  int64_t M, N, K;
  int32_t num_m_blocks;     // ceil(M / (128 * ncta))
  int32_t num_n_tiles;      // ceil(N / 256)
  int32_t tiles_per_group;  // most consecutive n-tiles of one work item (online softmax state lives across them)
  int32_t num_n_groups;     // work items per m-block: group g covers the n-tiles [g * tiles / groups, (g + 1) * tiles / groups)
                            // (balanced: sizes differ by at most one tile)
  int32_t rotate_groups;    // K1: m-blocks per wave of the persistent grid (0 = off).  The group a worker takes
                            // rotates from wave to wave, so that workers whose groups are one tile longer do not fall
                            // behind for the whole launch (CTAs that drift apart stop sharing W tiles and hidden
                            // panels in L2: profiles/r2_notes.md section 7)
  int32_t m_fast;           // item order: 0 = n-groups fastest (CTAs running together share A panels),
                            //             1 = m-blocks fastest (they share B panels)
  unsigned int* wave_sync;  // optional pair of global words {arrivals, abandoned} (zeroed before launch): the TMA
                            // producers of all CTAs rendezvous at every wave boundary of the static schedule, so
                            // that CTAs which stream the same operand panels stay inside each other's L2 window.
                            // A pure performance hint: if the grid is not fully co-resident (another kernel holds
                            // SMs) the first CTA whose wait exceeds kWaveSyncTimeout raises `abandoned` and every CTA
                            // runs on unsynchronised; nothing ever traps or deadlocks on it.
  // EPI_STATS
  const int64_t* targets;   // [M] global vocab ids
  int64_t v_offset;         // first vocab id of this slice
  float* parts;             // [num_n_groups, 3, M]: max, sum exp(z - max), target logit
  __nv_bfloat16* logits;    // optional [M, ld_logits]
  int64_t ld_logits;
  // exp-store mode (row_ref != nullptr, needs `logits`): what is stored is not the logit z but
  //   E[t, v] = exp(z[t, v] - row_ref[t])   (0 for rows with row_keep[t] == 0)
  // with a per-row reference known BEFORE the sweep, so that softmax = E * exp(row_ref - lse) is a per-row rescale and
  // the softmax backward folds into the epilogues / operands of the backward GEMMs by linearity (lmhead.cu).
  const float* row_ref;     // [M]
  const int32_t* row_keep;  // optional [M]
  // EPI_STORE / EPI_ACCUM
  void* out;                // [M, ld_out] bf16 or fp32
  int64_t ld_out;
  int32_t out_fp32;
  int32_t accumulate;
  // L2 eviction hints of the TMA traffic (0 = none, 1 = evict first, 2 = evict last)
  int32_t hint_a, hint_b, hint_store;
  // kXform: the A operand in HBM is the bf16 LOGITS chunk z[t, v]; the softmax backward
  //   P[t, v] = g[t] * ([v + v_offset == target[t]] - exp(z[t, v] - lse[t]))
  // is applied to every A tile in shared memory between its TMA load and the MMAs that read it
  // EPI_STORE fused with the reduce-scatter of the vocab-parallel path: row r of the output belongs to the rank that
  // owns token (row0 + r); the bf16 tile is stored STRAIGHT into that rank's peer-mapped slot buffer (NVLink P2P
  // stores from the epilogue, tile by tile while the GEMM runs), slot = this rank.  n_peers == 0: plain local store.
  __nv_bfloat16* peer_out[16];   // rank p's slot buffer [n_peers][slot_rows][ld_out]
  int32_t n_peers, my_rank;
  int64_t rows_per_owner, slot_rows, row0;
  // EPI_STORE with the softmax backward folded into the epilogue (exp-store path): the A operand is E = exp(z - ref),
  //   out[t, h] = a_t * acc[t, h] + g_t * W[tcol_t, h]      a_t = -g_t * exp(ref_t - lse_t)   (second term: target in slice)
  const int4* sm_rows;           // [M] SoftmaxEpiRow records, nullptr = plain store
  const __nv_bfloat16* sm_weight;   // W [K, N] row-major (the B operand), rows gathered at the target column
  int32_t prefetch_a;         // > 0: the producer prefetches the A tile of k-block kb + prefetch_a into L2
  const int4* x_rows;         // [tokens] packed per-token scalars (SoftmaxBwdRow, written by softmax_bwd_rows_kernel)
  int64_t x_tokens;           // number of token rows
};

// Per-token scalars of the exp-store backward (softmax_rows_kernel): P[t, v] = a * E[t, v] + g * [v == tcol].
struct SoftmaxEpiRow {
  float g;        // d loss / d logp (0: masked-out token)
  float a;        // -g * exp(row_ref - lse)
  int32_t tcol;   // target column within the slice, -1 if the target lives in another slice
  int32_t pad;
};

// Per-token scalars of the fused softmax backward, one 16-byte record so that a transform thread needs ONE load per
// token: P[t, v] = sgn(-g) * 2^(z * log2e + c) (+ g at column tcol of this vocabulary slice).
struct SoftmaxBwdRow {
  float g;        // d loss / d logp (0: masked-out token, the row becomes zeros)
  float c;        // log2|g| - lse * log2e
  int32_t tcol;   // target column within the slice, -1 if the target lives in another slice
  int32_t pad;
};

// kAcc = 1: 256-column tiles, the two TMEM accumulator stages double-buffer the epilogue.
// kAcc = 2: 512-column tiles (pairs only): both accumulators belong to ONE tile, every A smem tile
//           feeds twice the MMAs (L2->SM bytes per flop -25 %, half as many sweeps over the other
//           operand); the epilogue is not overlapped, fine when the K loop is hundreds of blocks.
__device__ __forceinline__ int item_n(const GemmParams& p, int item) {
  if (p.m_fast) return item / p.num_m_blocks;
  const int j = item % p.num_n_groups;
  if (p.rotate_groups <= 0) return j;
  return (j + (item / p.num_n_groups) / p.rotate_groups) % p.num_n_groups;
}
__device__ __forceinline__ void group_tiles(const GemmParams& p, int n_grp, int& t_begin, int& t_end) {
  t_begin = (int)(((int64_t)n_grp * p.num_n_tiles) / p.num_n_groups);
  t_end = (int)(((int64_t)(n_grp + 1) * p.num_n_tiles) / p.num_n_groups);
}
__device__ __forceinline__ int item_m(const GemmParams& p, int item) {
  return p.m_fast ? item % p.num_m_blocks : item / p.num_n_groups;
}

template <int kNCta, bool kStaging, int kAcc = 1>
struct GemmShape {
  static_assert(kAcc == 1 || (kAcc == 2 && kNCta == 2 && !kStaging), "512-column tiles need a CTA pair");
  static constexpr int BM = 128;                  // accumulator rows per CTA (TMEM lanes)
  static constexpr int UMMA_M = BM * kNCta;
  static constexpr int BN = 256;                  // UMMA_N = columns per accumulator
  static constexpr int TILE_N = BN * kAcc;        // output columns per work tile
  static constexpr int ACC_STAGES = (kAcc == 1) ? 2 : 1;
  static constexpr int BK = 64;                   // one 128-byte swizzle atom of bf16
  static constexpr int UMMA_K = 16;
  static constexpr int LOAD_BN = BN / kNCta;      // B rows per accumulator fetched by each CTA of the pair
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B1_BYTES = LOAD_BN * BK * 2;   // one accumulator's share of B
  static constexpr int B_BYTES = kAcc * B1_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (kNCta == 1) ? 4 : (kAcc == 1 ? 6 : 4);
  // K1 only: per epilogue warp 2 x [32 rows x 128 B] swizzled staging buffers for the TMA store of bf16 logits
  static constexpr int STAGING_BYTES = kStaging ? 4 * 2 * 4096 : 0;
  static constexpr int BAR_BYTES = (4 * STAGES + 4) * 8 + 16;   // full, empty, afull, ready per stage + 4 accumulator barriers
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;   // +1024: manual alignment
  static constexpr int TMEM_COLS = 512;
};

// longest legitimate wait at a wave rendezvous is the skew between CTAs of one wave (tens of microseconds);
// ~4 ms of SM clocks means that part of the grid is not resident
constexpr long long kWaveSyncTimeout = 8000000LL;
constexpr int kGemmThreads = 256;
constexpr int kGemmThreadsXform = 384;    // + warps 8..11: with warps 4..7 the eight transform warps of a kXform kernel
constexpr float kLog2e = 1.4426950408889634f;

// Half of one 128-byte row (32 of the 64 bf16 logits z of one token in a k-block; the row is 128B-swizzled: logical
// 16-byte chunk j sits at j ^ sw) rewritten in place to P = g * (onehot - exp(z - lse)):
//   P = -g * 2^(z * log2e - lse * log2e) = sgn(-g) * 2^(z * log2e + c),  c = log2|g| - lse * log2e
// so that an element costs one FFMA and one MUFU.EX2; the sign is XORed into the packed bf16 pairs, g is added at the
// target column `hot` (index within the row, -1 = not here).  A row with g == 0 (masked-out token) becomes zeros without
// being read, columns >= nvalid (outside the vocabulary slice, TMA zero fill) become zeros.  The four 16-byte loads
// are issued before any arithmetic (the volatile stores would otherwise serialise the chunks).
__device__ __forceinline__ void xform_half(uint32_t rowbase, uint32_t sw, int half, float g, float c, int hot, int nvalid) {
  const uint32_t j0 = (uint32_t)half * 4u;
  if (g == 0.f || nvalid <= (int)j0 * 8) {
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) ptx::st_shared_v4(rowbase + (((j0 + j) ^ sw) << 4), 0u, 0u, 0u, 0u);
    return;
  }
  const uint32_t sign = g > 0.f ? 0x80008000u : 0u;          // P has the sign of -g
  uint32_t w[4][4];
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    ptx::ld_shared_v4(rowbase + (((j0 + j) ^ sw) << 4), w[j][0], w[j][1], w[j][2], w[j][3]);
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    float f[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      f[2 * q] = ptx::ex2_approx(fmaf(__uint_as_float(w[j][q] << 16), kLog2e, c));
      f[2 * q + 1] = ptx::ex2_approx(fmaf(__uint_as_float(w[j][q] & 0xffff0000u), kLog2e, c));
    }
    const int col = (int)(j0 + j) * 8;                 // first column of this chunk within the row
    const bool fix = (hot >> 3) == (int)(j0 + j) || col + 8 > nvalid;
    if (fix) {                                         // rare: target column (once per token per sweep) / ragged edge
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v = g > 0.f ? -f[q] : f[q];
        if (col + q == hot) v += g;
        if (col + q >= nvalid) v = 0.f;
        f[q] = v;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
      w[j][q] = *reinterpret_cast<uint32_t*>(&t) ^ (fix ? 0u : sign);
    }
  }
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    ptx::st_shared_v4(rowbase + (((j0 + j) ^ sw) << 4), w[j][0], w[j][1], w[j][2], w[j][3]);
}

// kXform (256x512 pair tiles only): softmax-backward transform of the A tiles in shared memory, done by eight warps:
// the four epilogue warps (idle during the main loop of a kAcc = 2 tile) and four more (warps 8..11; one warp per SM
// sub-partition could not hide its own latencies: 2.6x slower than the MMAs it feeds).  The A tile lands on a
// CTA-LOCAL barrier (`afull`), each thread rewrites half a row in place, makes the writes visible to the async proxy
// and the warp arrives on the leader's `ready` barrier; the MMA issuer waits for `full` (B bytes of both CTAs) and
// `ready` (A of both CTAs transformed).
template <bool kAMN, bool kBMN, int kNCta, int kEpi, int kAcc = 1, bool kXform = false>
__global__ void __launch_bounds__(kXform ? kGemmThreadsXform : kGemmThreads, 1)
lmhead_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using S = GemmShape<kNCta, kEpi == EPI_STATS, kAcc>;
  constexpr int BM = S::BM, BN = S::BN, BK = S::BK, STAGES = S::STAGES;
  static_assert(!kXform || (kNCta == 2 && kAcc == 2 && kEpi != EPI_STATS), "the fused transform needs the idle epilogue warps of a 256x512 pair tile");

  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte aligned bases (in the shared window, identical in both CTAs of a pair)
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint8_t* staging = smem + STAGES * S::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* afull = bars + 2 * STAGES + 4;
  uint64_t* ready = bars + 3 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kNCta == 2) ? ptx::cluster_ctarank() : 0u;

  if (kNCta == 2) ptx::cluster_sync();
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (kEpi == EPI_STATS && p.logits != nullptr) ptx::prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full[i], kNCta);       // producer arrivals (leader's barrier collects both CTAs)
      ptx::mbar_init(&empty[i], 1);          // one tcgen05.commit
      ptx::mbar_init(&afull[i], 1);          // kXform: this CTA's A tile has landed
      ptx::mbar_init(&ready[i], kNCta * 8);  // kXform: one arrival per transform warp of each CTA
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);          // one tcgen05.commit
      ptx::mbar_init(&tempty[i], kNCta * 4); // one arrival per epilogue warp of each CTA
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kNCta>(tmem_slot, S::TMEM_COLS);
  ptx::tc_fence_before();
  if (kNCta == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  if (tmem_base != 0u) __trap();   // 512 columns = the whole TMEM of this SM (1 CTA/SM): base is lane 0, column 0

  const int num_workers = gridDim.x / kNCta;
  const int worker = blockIdx.x / kNCta;
  const int num_items = p.num_m_blocks * p.num_n_groups;
  const int num_k_blocks = (int)((p.K + BK - 1) / BK);

  // ---- kXform: softmax backward of one work tile's A operand, k-block by k-block, ahead of the MMAs (warps 4..11)
  auto xform_tile = [&](int m_blk, int& xstage, uint32_t& xphase) {
    const int xt = (int)threadIdx.x - 128;                                   // 0..255
    const int half = xt >> 7;
    auto load_row = [&](int64_t tok) -> int4 {                               // 16 bytes per token, L2 resident
      return tok < p.x_tokens ? __ldg(p.x_rows + tok) : make_int4(0, 0, -1, 0);
    };
    auto publish = [&]() {
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { if (rank == 0) ptx::mbar_arrive(&ready[xstage]); else ptx::mbar_arrive_cluster(&ready[xstage], 0); }
      if (++xstage == STAGES) { xstage = 0; xphase ^= 1u; }
    };
    if constexpr (!kAMN) {
      // K2a: A tile = [128 token rows][64 vocab columns], K-major: thread = (half, token row)
      const int r = xt & 127;
      const int4 rec = load_row((int64_t)m_blk * S::UMMA_M + (int64_t)rank * BM + r);
      const float g = __int_as_float(rec.x), c = __int_as_float(rec.y);
      const int64_t tcol = rec.z;
      const uint32_t sw = (uint32_t)r & 7u;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        ptx::mbar_wait(&afull[xstage], xphase);
        const uint32_t rowbase = ptx::smem_u32(smem_a + xstage * S::A_BYTES) + (uint32_t)r * 128u;
        const int64_t rel = tcol - (int64_t)kb * BK, left = p.K - (int64_t)kb * BK;
        xform_half(rowbase, sw, half, g, c, (rel >= 0 && rel < BK) ? (int)rel : -1, left < BK ? (int)left : BK);
        publish();
      }
    } else {
      // K2b: A tile = 2 atoms of [64 token rows][64 vocab columns], MN-major: thread = (half, atom, token row)
      const int xi = (xt >> 6) & 1, xk = xt & 63;
      const int64_t col0 = (int64_t)m_blk * S::UMMA_M + (int64_t)rank * BM + 64 * xi;      // first vocab column of the atom
      const int64_t left = p.M - col0;
      const int nvalid = left < 64 ? (int)left : 64;
      const uint32_t sw = (uint32_t)xk & 7u;
      int4 r0 = load_row(xk), r1 = load_row((int64_t)BK + xk);               // rows of k-blocks 0 and 1
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const float g0 = __int_as_float(r0.x), c0 = __int_as_float(r0.y);
        const int64_t rel = (int64_t)r0.z - col0;
        r0 = r1;
        r1 = load_row((int64_t)(kb + 2) * BK + xk);                         // two k-blocks ahead: in flight meanwhile
        ptx::mbar_wait(&afull[xstage], xphase);
        const uint32_t rowbase = ptx::smem_u32(smem_a + xstage * S::A_BYTES) + (uint32_t)xi * (BK * 128) + (uint32_t)xk * 128u;
        xform_half(rowbase, sw, half, g0, c0, (rel >= 0 && rel < 64) ? (int)rel : -1, nvalid);
        publish();
      }
    }
  };

  // The two single-instruction-stream roles run with the WHOLE warp converged and only the
  // TMA / tcgen05 instructions predicated on one elected lane: every operand is then provably
  // warp-uniform and lives in uniform registers (a lane-0-only branch makes ptxas wrap each
  // tcgen05.mma in an ELECT + R2UR.BROADCAST loop, ~130 instructions per k-block, which starved
  // the tensor pipe: see profiles/r1_notes.md).
  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0; uint32_t phase = 0;
    unsigned int sync_target = 0;
    bool sync_off = false;
    for (int item = worker; item < num_items; item += num_workers) {
      if (p.wave_sync != nullptr) {
        if (item != worker && !sync_off) {   // every CTA active in the previous wave has issued all of its loads
          bool off = false;
          if (ptx::elect_one()) {
            const long long t0 = clock64();
            while (ptx::ld_acquire_gpu(p.wave_sync) < sync_target) {
              if (ptx::ld_acquire_gpu(p.wave_sync + 1) != 0u) { off = true; break; }
              __nanosleep(200);
              if (clock64() - t0 > kWaveSyncTimeout) {          // peers are not resident: give the hint up for good
                ptx::red_release_gpu_add(p.wave_sync + 1, 1u);
                off = true;
                break;
              }
            }
          }
          sync_off = __any_sync(0xffffffffu, off);
        }
        const int wave_first = item - worker;                                  // first item of this wave
        sync_target += (unsigned int)(min(num_workers, num_items - wave_first) * kNCta);
      }
      const int n_grp = item_n(p, item), m_blk = item_m(p, item);
      const int m0 = m_blk * S::UMMA_M + (int)rank * BM;
      int t_begin, t_end;
      group_tiles(p, n_grp, t_begin, t_end);
      for (int nt = t_begin; nt < t_end; ++nt) {
        const int n0 = nt * S::TILE_N + (int)rank * S::LOAD_BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u);
          if (ptx::elect_one()) {
            const int k0 = kb * BK;
            uint8_t* sa = smem_a + stage * S::A_BYTES;
            uint8_t* sb = smem_b + stage * S::B_BYTES;
            if constexpr (kXform) {
              ptx::mbar_expect_tx(&afull[stage], S::A_BYTES);                    // local: the transform warps wait here
              if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * S::B_BYTES);
              else ptx::mbar_arrive_cluster(&full[stage], 0);
            } else if constexpr (kNCta == 1) {
              ptx::mbar_expect_tx(&full[stage], S::STAGE_BYTES);
            } else {
              if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * S::STAGE_BYTES);
              else ptx::mbar_arrive_cluster(&full[stage], 0);
            }
            auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
              const int hint = (tm == &tmA) ? p.hint_a : p.hint_b;
              if (hint == 0) {
                if constexpr (kNCta == 1) ptx::tma_load_2d(dst, tm, &full[stage], c0, c1);
                else ptx::tma_load_2d_2sm(dst, tm, &full[stage], c0, c1);
              } else {
                if constexpr (kNCta == 1) ptx::tma_load_2d_hint(dst, tm, &full[stage], c0, c1, ptx::l2_policy(hint));
                else ptx::tma_load_2d_2sm_hint(dst, tm, &full[stage], c0, c1, ptx::l2_policy(hint));
              }
            };
            if (p.prefetch_a > 0 && kb + p.prefetch_a < num_k_blocks) {   // A streams from HBM: warm L2 ahead
              const int kp = (kb + p.prefetch_a) * BK;
              if constexpr (!kAMN) {
                ptx::tma_prefetch_l2_2d(&tmA, kp, m0);
              } else {
#pragma unroll
                for (int i = 0; i < BM / 64; ++i) ptx::tma_prefetch_l2_2d(&tmA, m0 + 64 * i, kp);
              }
            }
            if constexpr (kXform) {                                     // plain (1-SM) TMA onto the local barrier
              if constexpr (!kAMN) {
                ptx::tma_load_2d(sa, &tmA, &afull[stage], k0, m0);
              } else {
#pragma unroll
                for (int i = 0; i < BM / 64; ++i) ptx::tma_load_2d(sa + i * (BK * 128), &tmA, &afull[stage], m0 + 64 * i, k0);
              }
            } else if constexpr (!kAMN) {
              load(sa, &tmA, k0, m0);                                   // [128 rows][64 k] K-major
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) load(sa + i * (BK * 128), &tmA, m0 + 64 * i, k0);   // [64 k][64 m] atoms
            }
#pragma unroll
            for (int j = 0; j < kAcc; ++j) {                 // one B slab per accumulator
              uint8_t* sbj = sb + j * S::B1_BYTES;
              if constexpr (!kBMN) {
                load(sbj, &tmB, k0, n0 + j * BN);
              } else {
#pragma unroll
                for (int i = 0; i < S::LOAD_BN / 64; ++i) load(sbj + i * (BK * 128), &tmB, n0 + j * BN + 64 * i, k0);
              }
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.wave_sync != nullptr) {
        if (ptx::elect_one()) ptx::red_release_gpu_add(p.wave_sync, 1u);
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(S::UMMA_M, BN, kAMN, kBMN);
      // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO), one atom along K (LBO unused)
      // MN-major SW128: 64-element atoms along MN 8 KB apart (LBO), 8-k groups 1024 B apart (SBO)
      constexpr uint32_t a_lbo = kAMN ? BK * 128 : 0, b_lbo = kBMN ? BK * 128 : 0;
      constexpr uint32_t a_kstep = kAMN ? S::UMMA_K * 128 : S::UMMA_K * 2;
      constexpr uint32_t b_kstep = kBMN ? S::UMMA_K * 128 : S::UMMA_K * 2;
      // descriptors of stage 0 / k 0; later ones differ only in the 14-bit start-address field
      const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), a_lbo, 1024);
      const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), b_lbo, 1024);
      int stage = 0; uint32_t phase = 0; uint32_t acc_iter = 0;
      for (int item = worker; item < num_items; item += num_workers) {
        const int n_grp = item_n(p, item);
        int t_begin, t_end;
        group_tiles(p, n_grp, t_begin, t_end);
        for (int nt = t_begin; nt < t_end; ++nt, ++acc_iter) {
          const uint32_t a = acc_iter % S::ACC_STAGES, aphase = (acc_iter / S::ACC_STAGES) & 1u;
          ptx::mbar_wait(&tempty[a], aphase ^ 1u);           // epilogue has drained this accumulator
          ptx::tc_fence_after();
          const uint32_t tmem_d = a * S::TILE_N;             // TMEM base is 0: this CTA owns all 512 columns
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            ptx::mbar_wait(&full[stage], phase);             // TMA bytes have landed (both CTAs)
            if constexpr (kXform) ptx::mbar_wait(&ready[stage], phase);   // ... and both A tiles are rewritten
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint64_t da = da0 + (uint64_t)((uint32_t)(stage * S::A_BYTES) >> 4);
              const uint64_t db = db0 + (uint64_t)((uint32_t)(stage * S::B_BYTES) >> 4);
#pragma unroll
              for (int j = 0; j < kAcc; ++j)
#pragma unroll
                for (int k = 0; k < BK / S::UMMA_K; ++k)
                  ptx::umma_bf16<kNCta>(tmem_d + j * BN, da + (uint64_t)((k * a_kstep) >> 4),
                                        db + (uint64_t)((j * S::B1_BYTES + k * b_kstep) >> 4), idesc,
                                        (kb | k) != 0 ? 1u : 0u);
              ptx::umma_commit<kNCta>(&empty[stage]);        // frees the smem stage when the MMAs retire
              if (kb == num_k_blocks - 1) ptx::umma_commit<kNCta>(&tfull[a]);   // accumulator ready
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (kNCta == 2 && acc_iter > 0) {
        // the peer's last remote arrivals must land before this CTA may exit
        const uint32_t last = acc_iter - 1;
        ptx::mbar_wait(&tempty[last % S::ACC_STAGES], (last / S::ACC_STAGES) & 1u);
      }
    }
  } else if (kXform && warp >= 8) {
    // ================================ extra transform warps ================================
    int xstage = 0; uint32_t xphase = 0;
    for (int item = worker; item < num_items; item += num_workers) {
      const int n_grp = item_n(p, item), m_blk = item_m(p, item);
      int t_begin, t_end;
      group_tiles(p, n_grp, t_begin, t_end);
      for (int nt = t_begin; nt < t_end; ++nt) xform_tile(m_blk, xstage, xphase);
    }
  } else if (warp >= 4) {
    // ================================ epilogue warps ================================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row_in_tile = q * 32 + lane;
    const uint32_t tmem_row = ((uint32_t)(q * 32) << 16);
    uint32_t acc_iter = 0;
    const uint32_t stg_warp = ptx::smem_u32(staging) + (uint32_t)(warp - 4) * 8192u;   // this warp's 2 buffers
    uint32_t sbuf = 0;
    int xstage = 0; uint32_t xphase = 0;          // kXform: position in the smem ring (same sequence as the producer)
    for (int item = worker; item < num_items; item += num_workers) {
      const int n_grp = item_n(p, item), m_blk = item_m(p, item);
      const int64_t row = (int64_t)m_blk * S::UMMA_M + (int64_t)rank * BM + row_in_tile;
      const bool row_ok = row < p.M;
      int t_begin, t_end;
      group_tiles(p, n_grp, t_begin, t_end);

      float run_max = -INFINITY, run_sum = 0.f, tgt_logit = 0.f;
      int64_t tgt_col = -1;
      float sm_g = 0.f, sm_a = 0.f;                          // EPI_STORE softmax epilogue: this row's record
      int sm_tcol = -1;
      if constexpr (kEpi == EPI_STORE) {
        if (p.sm_rows != nullptr && row_ok) {
          const int4 rec = __ldg(p.sm_rows + row);
          sm_g = __int_as_float(rec.x); sm_a = __int_as_float(rec.y); sm_tcol = rec.z;
        }
      }
      float ref_l2e = 0.f;                                   // exp-store mode: row_ref[row] * log2e
      bool keep_row = false;
      if constexpr (kEpi == EPI_STATS) {
        if (row_ok) tgt_col = p.targets[row] - p.v_offset;   // may fall outside this slice
        if (tgt_col >= p.N) tgt_col = -1;                    // (never match a zero-filled column)
        if (p.row_ref != nullptr && row_ok) {
          ref_l2e = p.row_ref[row] * kLog2e;
          keep_row = p.row_keep == nullptr || p.row_keep[row] != 0;
        }
      }

      for (int nt = t_begin; nt < t_end; ++nt, ++acc_iter) {
        const uint32_t a = acc_iter % S::ACC_STAGES, aphase = (acc_iter / S::ACC_STAGES) & 1u;
        if constexpr (kXform) xform_tile(m_blk, xstage, xphase);
        const int64_t n0 = (int64_t)nt * S::TILE_N;
        const int64_t n_left = p.N - n0;
        const int n_valid = n_left < (int64_t)S::TILE_N ? (int)n_left : S::TILE_N;          // columns of this tile inside N
        // EPI_STORE with the softmax epilogue: this row's segment of W[tcol] is gathered one 32-column chunk AHEAD of
        // its use (a dependent L2 / HBM load per chunk inside `process` cost ~1 us x 16 chunks per tile: 7 % of K2a
        // when the vocabulary slice, i.e. the K loop, is 1/8 of the full one); the segment is pulled into L2 and its
        // first chunk loaded while the tile's MMAs still run
        uint4 wa[4] = {}, wb[4] = {};
        auto gather_w = [&](uint4 (&wq)[4], const int c) {
          if constexpr (kEpi == EPI_STORE) {
            if (sm_tcol >= 0 && c * 32 < n_valid) {
              const uint4* wrow = reinterpret_cast<const uint4*>(p.sm_weight + (int64_t)sm_tcol * p.N + n0 + c * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c * 32 + j * 8 < n_valid) wq[j] = __ldg(wrow + j);
            }
          }
        };
        if constexpr (kEpi == EPI_STORE) {
          if (sm_tcol >= 0) {
            const __nv_bfloat16* wseg = p.sm_weight + (int64_t)sm_tcol * p.N + n0;
            for (int cc = 64; cc < n_valid; cc += 64) ptx::prefetch_l2(wseg + cc);       // 128-byte lines
          }
          gather_w(wa, 0);
        }
        ptx::mbar_wait(&tfull[a], aphase);
        ptx::tc_fence_after();
        const int tgt_in_tile = (tgt_col >= n0 && tgt_col < n0 + S::TILE_N) ? (int)(tgt_col - n0) : -1;
        const uint32_t tmem_tile = tmem_row + a * S::TILE_N;

        // bf16 copy of one 32-column chunk (scaled by `scale`) for the chunked backward: [32 rows x 64 cols] per warp
        // staged in 128B-swizzled smem (conflict-free 16-byte stores), then ONE coalesced TMA store per 64 columns;
        // the box is clipped against [T, V] by the TMA unit, so ragged edges need no guards.
        auto store_chunk = [&](const uint32_t (&v)[32], const int c, const bool scaled, const float scale_in) {
          const float scale = scaled ? scale_in : 1.0f;
          const int cbase = c * 32;
          const int half = c & 1;
          if (half == 0) {
            if (lane == 0) ptx::bulk_wait_read<1>();   // the store that last read this buffer is done
            __syncwarp();
          }
          const uint32_t rowbase = stg_warp + sbuf * 4096u + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 t0 = __floats2bfloat162_rn(scaled ? __uint_as_float(v[j * 8 + 0]) * scale : __uint_as_float(v[j * 8 + 0]), scaled ? __uint_as_float(v[j * 8 + 1]) * scale : __uint_as_float(v[j * 8 + 1]));
            __nv_bfloat162 t1 = __floats2bfloat162_rn(scaled ? __uint_as_float(v[j * 8 + 2]) * scale : __uint_as_float(v[j * 8 + 2]), scaled ? __uint_as_float(v[j * 8 + 3]) * scale : __uint_as_float(v[j * 8 + 3]));
            __nv_bfloat162 t2 = __floats2bfloat162_rn(scaled ? __uint_as_float(v[j * 8 + 4]) * scale : __uint_as_float(v[j * 8 + 4]), scaled ? __uint_as_float(v[j * 8 + 5]) * scale : __uint_as_float(v[j * 8 + 5]));
            __nv_bfloat162 t3 = __floats2bfloat162_rn(scaled ? __uint_as_float(v[j * 8 + 6]) * scale : __uint_as_float(v[j * 8 + 6]), scaled ? __uint_as_float(v[j * 8 + 7]) * scale : __uint_as_float(v[j * 8 + 7]));
            const uint32_t chunk16 = (uint32_t)(half * 4 + j) ^ ((uint32_t)lane & 7u);   // 128B swizzle
            ptx::st_shared_v4(rowbase + (chunk16 << 4), *reinterpret_cast<uint32_t*>(&t0),
                              *reinterpret_cast<uint32_t*>(&t1), *reinterpret_cast<uint32_t*>(&t2),
                              *reinterpret_cast<uint32_t*>(&t3));
          }
          if (half == 1) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int32_t r0 = (int32_t)((int64_t)m_blk * S::UMMA_M + (int64_t)rank * BM + q * 32);
              if (p.hint_store == 0) ptx::tma_store_2d(&tmC, stg_warp + sbuf * 4096u, (int32_t)(n0 + cbase - 32), r0);
              else ptx::tma_store_2d_hint(&tmC, stg_warp + sbuf * 4096u, (int32_t)(n0 + cbase - 32), r0,
                                          ptx::l2_policy(p.hint_store));
              ptx::bulk_commit();
            }
            sbuf ^= 1u;
          }
        };

        // one 32-column chunk of this thread's accumulator row
        auto process = [&](uint32_t (&v)[32], const int c, const uint4 (&wq)[4]) {
          const int cbase = c * 32;                       // first column of the chunk within the tile
          if constexpr (kEpi == EPI_STATS) {
            if (p.logits != nullptr && p.row_ref == nullptr) store_chunk(v, c, false, 1.0f);     // bf16 logits (dlogits path)
            if (cbase + 32 > n_valid) {            // ragged last tile: TMA zero-filled columns are not vocabulary
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j >= n_valid) v[j] = __float_as_uint(-INFINITY);
            }
            if ((tgt_in_tile >> 5) == c) {          // rare: once per row over the whole sweep (-1 >> 5 == -1)
              const int jj = tgt_in_tile & 31;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j == jj) tgt_logit = __uint_as_float(v[j]);
            }
            float cmax = __uint_as_float(v[0]);
#pragma unroll
            for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, __uint_as_float(v[j]));
            const float new_max = fmaxf(run_max, cmax);
            if (new_max != -INFINITY) {            // (an all-masked chunk before any valid column cannot occur)
              const float neg = -new_max * kLog2e;
              float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;      // 4 chains: the adds do not serialise
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float f0 = ptx::ex2_approx(fmaf(__uint_as_float(v[j + 0]), kLog2e, neg));
                const float f1 = ptx::ex2_approx(fmaf(__uint_as_float(v[j + 1]), kLog2e, neg));
                const float f2 = ptx::ex2_approx(fmaf(__uint_as_float(v[j + 2]), kLog2e, neg));
                const float f3 = ptx::ex2_approx(fmaf(__uint_as_float(v[j + 3]), kLog2e, neg));
                acc0 += f0; acc1 += f1; acc2 += f2; acc3 += f3;
                // exp-store mode: the exponentials themselves are what the backward needs (dead stores otherwise)
                v[j + 0] = __float_as_uint(f0); v[j + 1] = __float_as_uint(f1);
                v[j + 2] = __float_as_uint(f2); v[j + 3] = __float_as_uint(f3);
              }
              if (p.row_ref != nullptr)      // E = exp(z - max) * exp(max - ref); masked-out rows store zeros
                store_chunk(v, c, true, keep_row ? ptx::ex2_approx(fmaf(new_max, kLog2e, -ref_l2e)) : 0.f);
              // rescale from the EXACT difference: an unchanged max must give a factor of exactly 1
              // (via fmaf(run_max, log2e, neg) the rounding of max*log2e would compound over the
              // ~4.7k chunks of a vocabulary sweep)
              run_sum = run_sum * ptx::ex2_approx((run_max - new_max) * kLog2e) + ((acc0 + acc1) + (acc2 + acc3));
              run_max = new_max;
            }
          } else if constexpr (kEpi == EPI_STORE) {
            if (p.sm_rows != nullptr && row_ok) {          // softmax backward by linearity: rescale + target row of W
              if (sm_g == 0.f) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;    // masked-out token: exactly zero (never 0 * inf)
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(sm_a * __uint_as_float(v[j]));
                if (sm_tcol >= 0 && cbase < n_valid) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    if (cbase + j * 8 < n_valid) {         // N % 8 == 0: whole 16-byte groups
                      const uint4 w8 = wq[j];              // W[tcol, n0 + cbase + 8 j ..], loaded one chunk ahead
                      const uint32_t ww[4] = {w8.x, w8.y, w8.z, w8.w};
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        v[j * 8 + 2 * k] = __float_as_uint(fmaf(sm_g, __uint_as_float(ww[k] << 16), __uint_as_float(v[j * 8 + 2 * k])));
                        v[j * 8 + 2 * k + 1] = __float_as_uint(fmaf(sm_g, __uint_as_float(ww[k] & 0xffff0000u), __uint_as_float(v[j * 8 + 2 * k + 1])));
                      }
                    }
                  }
                }
              }
            }
            if (row_ok) {
              if (p.out_fp32) {
                float* dst = reinterpret_cast<float*>(p.out) + row * p.ld_out + n0 + cbase;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (cbase + j * 4 < n_valid)
                    *reinterpret_cast<uint4*>(dst + j * 4) = make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
              } else {
                __nv_bfloat16* dst;
                if (p.n_peers > 0) {
                  const int64_t grow = p.row0 + row;                       // global token row
                  int owner = (int)(grow / p.rows_per_owner);
                  if (owner >= p.n_peers) owner = p.n_peers - 1;
                  dst = p.peer_out[owner] + ((int64_t)p.my_rank * p.slot_rows + (grow - (int64_t)owner * p.rows_per_owner)) * p.ld_out + n0 + cbase;
                } else {
                  dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + n0 + cbase;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (cbase + j * 8 < n_valid) {
                    uint4 pk;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                    __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                    __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                    pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                    pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                    *reinterpret_cast<uint4*>(dst + j * 8) = pk;
                  }
                }
              }
            }
          } else {  // EPI_ACCUM: fp32 store, or fire-and-forget 16-byte reductions (one writer per element
                    // per launch, so the result is still deterministic; no read latency in the epilogue)
            if (row_ok) {
              float* dst = reinterpret_cast<float*>(p.out) + row * p.ld_out + n0 + cbase;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (cbase + j * 4 < n_valid) {
                  if (p.accumulate)
                    ptx::red_add_v4_f32(dst + j * 4, __uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]),
                                        __uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3]));
                  else
                    *reinterpret_cast<uint4*>(dst + j * 4) = make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                }
              }
            }
          }
        };

        // software pipeline: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
        constexpr int kChunks = S::TILE_N / 32;
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32b_x32(tmem_tile, va);
#pragma unroll 1
        for (int c = 0; c < kChunks; c += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(tmem_tile + (c + 1) * 32, vb);
          gather_w(wb, c + 1);
          process(va, c, wa);
          ptx::tmem_ld_wait();
          if (c + 2 < kChunks) {
            ptx::tmem_ld_32x32b_x32(tmem_tile + (c + 2) * 32, va);
            gather_w(wa, c + 2);
          }
          process(vb, c + 1, wb);
        }
        // all tcgen05.ld of this warp have completed (wait::ld above): hand the accumulator back
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kNCta == 1) ptx::mbar_arrive(&tempty[a]);
          else ptx::mbar_arrive_cluster(&tempty[a], 0);
        }
      }

      if constexpr (kEpi == EPI_STATS) {
        if (row_ok) {
          float* part = p.parts + (int64_t)n_grp * 3 * p.M;
          part[row] = run_max;
          part[p.M + row] = run_sum;
          part[2 * p.M + row] = tgt_logit;
        }
      }
    }
  }

  if (kEpi == EPI_STATS && warp >= 4 && lane == 0 && p.logits != nullptr) ptx::bulk_wait<0>();   // smem must outlive the stores

  // ================================ teardown ================================
  ptx::tc_fence_before();
  if (kNCta == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kNCta>(tmem_base, S::TMEM_COLS);
  }
}

}  // namespace o3v
