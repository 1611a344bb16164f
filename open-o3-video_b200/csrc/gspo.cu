// K3a (first-EOS mask) and K3 (KL + group advantages + GSPO ratio/clip/loss, fwd+bwd).
//
// Replaces trainer/grpo_trainer.py:590-596 and :635-636, 658, 675-681, 691-706, 711, 737
// of the reference.  Both are launch / latency bound (about 20 B per token, a few MB per step).  K3 is ONE
// launch over a (slice, sequence) grid: every CTA reduces the masked partial sums of its slice with warp
// shuffles; the CTA that arrives LAST for a sequence (per-sequence ticket) adds the slices up in fixed
// slice order, writes the per-sequence outputs and d loss / d logp of the whole sequence (its inputs are
// L2 hits by then), and the CTA that finishes the last sequence of the step takes the mean over the
// sequences in index order.  No CTA ever waits for another one, and results are run-to-run bit-stable.
#include <algorithm>

#include "common.cuh"

namespace o3v {

constexpr int kGspoThreads = 256;

// ------------------------------------------------------------------------------------
// K3a: eos_idx[n] = first t with ids[n,t]==eos else Tc; mask[n,t] = (t <= eos_idx[n]).
// ------------------------------------------------------------------------------------
constexpr int kEosThreads = 512;
constexpr int kEosBatch = 8;     // independent 8-byte loads in flight per thread

__global__ void __launch_bounds__(kEosThreads)
eos_mask_kernel(const int64_t* __restrict__ ids, int64_t Tc, int64_t eos_id,
                int64_t* __restrict__ eos_idx, int32_t* __restrict__ mask) {
  __shared__ int s_min[kEosThreads / 32];
  const int64_t n = blockIdx.x;
  const int64_t* row = ids + n * Tc;
  int first = (int)Tc;
  // strided scan in batches of kEosBatch loads issued together (one memory round trip per batch instead of one
  // per element); a thread's indices ascend, so its first hit is its minimum and it can stop there
  for (int64_t t0 = threadIdx.x; t0 < Tc && first == (int)Tc; t0 += (int64_t)kEosThreads * kEosBatch) {
    int64_t v[kEosBatch];
#pragma unroll
    for (int j = 0; j < kEosBatch; ++j) {
      const int64_t t = t0 + (int64_t)j * kEosThreads;
      v[j] = t < Tc ? row[t] : eos_id - 1;          // (eos_id - 1 != eos_id: never a hit)
    }
#pragma unroll
    for (int j = kEosBatch - 1; j >= 0; --j) {
      const int64_t t = t0 + (int64_t)j * kEosThreads;
      if (t < Tc && v[j] == eos_id) first = (int)t;  // descending j: the smallest index wins
    }
  }
  first = warp_min_i(first);
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = first;
  __syncthreads();
  first = s_min[0];
#pragma unroll
  for (int w = 1; w < kEosThreads / 32; ++w) first = min(first, s_min[w]);
  if (threadIdx.x == 0) eos_idx[n] = (int64_t)first;
  int32_t* mrow = mask + n * Tc;
  for (int64_t t = threadIdx.x; t < Tc; t += kEosThreads) mrow[t] = (t <= (int64_t)first) ? 1 : 0;
}

// ------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------
constexpr int kMaxSlices = 32;   // CTAs per sequence (upper bound; fixes the workspace layout)

struct GspoParams {
  const float* logp; const float* old_logp; const float* ref; const int32_t* mask;
  const float* rpf;
  int64_t N, Tc, F, G;      // N = TOTAL sequences of the step (loss is a mean over N)
  int64_t seq_offset;       // this launch handles global sequences [seq_offset, seq_offset + gridDim.y)
  int32_t S; int32_t slice_len;   // CTAs per sequence, tokens per CTA
  float beta, eps_lo, eps_hi; int gspo;
  float* loss; float* mean_kl; float* adv; float* rstd; int32_t* clen;
  float* grad; float* kl_out;
  float* seq_loss; float* seq_kl; float* partials; unsigned int* ticket; unsigned int* seq_ticket;   // workspace
};

__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kGspoThreads / 32; ++w) r += smem[w];   // fixed order: deterministic
  return r;
}

// KL term of grpo_trainer.py:635-636 and its derivative w.r.t. logp.
__device__ __forceinline__ void kl_and_grad(float lp, float rf, float& kl, float& dkl) {
  const float x = rf - lp;
  const float xc = fminf(fmaxf(x, -10.f), 10.f);
  const float e = expf(xc);
  kl = e - xc - 1.f;
  // clamp passes gradient on the closed interval [-10, 10]; d(e^x - x - 1)/dlogp = 1 - e^x
  dkl = (x >= -10.f && x <= 10.f) ? (1.f - e) : 0.f;
}

// group advantage of global sequence n (grpo_trainer.py:658, 675-681); G is small, one thread does it
__device__ __forceinline__ void group_advantage(const GspoParams& p, int64_t n, float& adv, float& sd) {
  const int64_t g0 = (n / p.G) * p.G;
  float mine = 0.f, mean = 0.f;
  for (int64_t j = 0; j < p.G; ++j) {
    float r = 0.f;
    for (int64_t f = 0; f < p.F; ++f) r += p.rpf[(g0 + j) * p.F + f];
    if (g0 + j == n) mine = r;
    mean += r;
  }
  mean /= (float)p.G;
  float var = 0.f;
  for (int64_t j = 0; j < p.G; ++j) {
    float r = 0.f;
    for (int64_t f = 0; f < p.F; ++f) r += p.rpf[(g0 + j) * p.F + f];
    var += (r - mean) * (r - mean);
  }
  sd = sqrtf(var / (float)(p.G - 1));   // unbiased; G == 1 -> NaN like torch.std
  adv = (mine - mean) / (sd + 1e-4f);
}

// Same, computed by the whole CTA: thread j < G sums the F rewards of rollout j of the group (independent loads) into
// shared memory, thread 0 reduces the G values in index order (same arithmetic and order as group_advantage).
constexpr int kMaxGroupSmem = 256;
__device__ __forceinline__ void group_advantage_cta(const GspoParams& p, int64_t n, float* s_r, float& adv, float& sd) {
  const int64_t g0 = (n / p.G) * p.G;
  if (p.G > kMaxGroupSmem) {                     // (never in practice: G = 4 .. 16)
    if (threadIdx.x == 0) { group_advantage(p, n, adv, sd); s_r[0] = adv; s_r[1] = sd; }
    __syncthreads();
    adv = s_r[0]; sd = s_r[1];
    __syncthreads();
    return;
  }
  for (int64_t j = threadIdx.x; j < p.G; j += kGspoThreads) {
    float r = 0.f;
    for (int64_t f = 0; f < p.F; ++f) r += p.rpf[(g0 + j) * p.F + f];
    s_r[j] = r;
  }
  __syncthreads();
  float mean = 0.f;
  for (int64_t j = 0; j < p.G; ++j) mean += s_r[j];
  mean /= (float)p.G;
  float var = 0.f;
  for (int64_t j = 0; j < p.G; ++j) var += (s_r[j] - mean) * (s_r[j] - mean);
  sd = sqrtf(var / (float)(p.G - 1));            // unbiased; G == 1 -> NaN like torch.std
  adv = (s_r[n - g0] - mean) / (sd + 1e-4f);
  __syncthreads();
}

// One launch: masked partial sums of (slice, sequence); the last CTA to arrive for a sequence finishes it.
__global__ void __launch_bounds__(kGspoThreads)
gspo_kernel(const GspoParams p) {
  __shared__ float s_red[kGspoThreads / 32];
  __shared__ float s_tot[8];
  __shared__ float s_r[kMaxGroupSmem];
  __shared__ bool s_last;
  const int64_t nl = blockIdx.y, n = p.seq_offset + nl;
  float A_seq, sd_seq;
  group_advantage_cta(p, n, s_r, A_seq, sd_seq);
  const int64_t Tc = p.Tc;
  const float* __restrict__ lp_row = p.logp + nl * Tc;
  const float* __restrict__ old_row = p.old_logp ? p.old_logp + nl * Tc : nullptr;
  const float* __restrict__ ref_row = p.ref + nl * Tc;
  const int32_t* __restrict__ m_row = p.mask + nl * Tc;
  float* __restrict__ kl_row = p.kl_out ? p.kl_out + nl * Tc : nullptr;
  {
    // ---- this CTA's slice: masked partial sums -> partials[n][slice][4]
    const int64_t t0 = (int64_t)blockIdx.x * p.slice_len, t1 = min(t0 + p.slice_len, Tc);
    const float A = A_seq;
    float cnt = 0.f, sum_lr = 0.f, sum_kl = 0.f, sum_obj = 0.f;
#pragma unroll 4
    for (int64_t t = t0 + threadIdx.x; t < t1; t += kGspoThreads) {
      const float lp = lp_row[t], rf = ref_row[t];
      const float m = (float)m_row[t];
      float kl, dkl;
      kl_and_grad(lp, rf, kl, dkl);
      if (kl_row) kl_row[t] = kl;
      const float lr = old_row ? (lp - old_row[t]) : 0.f;     // :691 (x - x.detach() == 0 exactly)
      cnt += m;
      sum_lr += lr * m;
      sum_kl += kl * m;
      if (!p.gspo) {                                          // token-level branch :696
        const float c1 = expf(lr);
        const float c2 = fminf(fmaxf(c1, 1.f - p.eps_lo), 1.f + p.eps_hi);
        sum_obj += -fminf(c1 * A, c2 * A) * m;
      }
    }
    cnt = block_sum(cnt, s_red);
    sum_lr = block_sum(sum_lr, s_red);
    sum_kl = block_sum(sum_kl, s_red);
    sum_obj = block_sum(sum_obj, s_red);
    if (threadIdx.x == 0) {
      float* part = p.partials + (n * kMaxSlices + blockIdx.x) * 4;
      __stcg(part + 0, cnt); __stcg(part + 1, sum_lr); __stcg(part + 2, sum_kl); __stcg(part + 3, sum_obj);
      __threadfence();
      const unsigned int arrived = atomicAdd(p.seq_ticket + n, 1u);
      s_last = (arrived == (unsigned int)(p.S - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
  }

  // ---- last CTA of sequence n: totals in fixed slice order, per-sequence outputs, d loss / d logp
  if (threadIdx.x == 0) {
    float cnt = 0.f, sum_lr = 0.f, sum_kl = 0.f, sum_obj = 0.f;
    for (int sidx = 0; sidx < p.S; ++sidx) {
      const float* part = p.partials + (n * kMaxSlices + sidx) * 4;
      cnt += __ldcg(part + 0); sum_lr += __ldcg(part + 1); sum_kl += __ldcg(part + 2); sum_obj += __ldcg(part + 3);
    }
    const float A = A_seq, sd = sd_seq;
    const float denom = fmaxf(cnt, 1.f);                      // .clamp(min=1.0) :693, :706
    float c1_seq = 1.f, gate_seq = 1.f;
    if (p.gspo) {
      const float sl = sum_lr / denom;                        // :693
      c1_seq = expf(sl);                                      // :698
      const float c2 = fminf(fmaxf(c1_seq, 1.f - p.eps_lo), 1.f + p.eps_hi);   // :699
      sum_obj = -fminf(c1_seq * A, c2 * A) * cnt;             // :701-703 (constant over t)
      // d(-min(c1 A, clamp(c1) A))/dc1 = -A * gate  (torch.minimum splits ties, clamp passes
      // gradient on the closed interval, so inside the clip range the two halves add up)
      gate_seq = (A > 0.f) ? (c1_seq <= 1.f + p.eps_hi ? 1.f : 0.f)
               : (A < 0.f) ? (c1_seq >= 1.f - p.eps_lo ? 1.f : 0.f) : 1.f;
    }
    p.seq_loss[n] = (sum_obj + p.beta * sum_kl) / denom;      // :704-706
    p.seq_kl[n] = sum_kl / cnt;                               // :737 (no clamp: 0/0 -> NaN as reference)
    if (p.clen) p.clen[n] = (int32_t)cnt;                     // :711
    if (p.adv) p.adv[n] = A;
    if (p.rstd) p.rstd[n] = sd;
    p.seq_ticket[n] = 0u;                                     // re-armed
    s_tot[0] = A; s_tot[1] = cnt; s_tot[2] = denom; s_tot[3] = c1_seq; s_tot[4] = gate_seq;
  }
  __syncthreads();
  if (p.grad) {
    const float A = s_tot[0], cnt = s_tot[1], denom = s_tot[2], c1_seq = s_tot[3], gate_seq = s_tot[4];
    const float scale = 1.f / (denom * (float)p.N);
    // GSPO: the per-token objective is constant over t, its masked mean is obj*cnt/denom,
    // and ds/dlogp_t = mask_t/denom, so the factor is (cnt/denom) * mask_t/denom.
    const float w = p.gspo ? (cnt / denom) : 1.f;
    float* __restrict__ g_row = p.grad + nl * Tc;
    auto grad_of = [&](float lp, float rf, float mf, float ol) -> float {
      float kl, dkl;
      kl_and_grad(lp, rf, kl, dkl);
      float dobj;
      if (p.gspo) {
        dobj = -A * c1_seq * gate_seq;
      } else {
        const float lr = old_row ? (lp - ol) : 0.f;
        const float c1 = expf(lr);
        const float gate = (A > 0.f) ? (c1 <= 1.f + p.eps_hi ? 1.f : 0.f)
                         : (A < 0.f) ? (c1 >= 1.f - p.eps_lo ? 1.f : 0.f) : 1.f;
        dobj = -A * c1 * gate;
      }
      return mf * scale * (dobj * w + p.beta * dkl);
    };
    // one CTA walks the whole sequence (its inputs are L2 hits by now): 16-byte vectors, two independent vectors per
    // thread and iteration, so a 16384-token sequence is 8 memory round trips per thread instead of 64
    const bool vec = ((Tc & 3) == 0) &&
                     ((((uintptr_t)lp_row | (uintptr_t)ref_row | (uintptr_t)m_row | (uintptr_t)g_row |
                        (uintptr_t)(old_row ? old_row : lp_row)) & 15u) == 0);
    if (vec) {
      const int64_t nv = Tc >> 2;
      const float4* lp4 = reinterpret_cast<const float4*>(lp_row);
      const float4* rf4 = reinterpret_cast<const float4*>(ref_row);
      const int4* m4 = reinterpret_cast<const int4*>(m_row);
      const float4* ol4 = reinterpret_cast<const float4*>(old_row ? old_row : lp_row);
      float4* g4 = reinterpret_cast<float4*>(g_row);
#pragma unroll 2
      for (int64_t i = threadIdx.x; i < nv; i += kGspoThreads) {
        const float4 a = lp4[i], b = rf4[i], o = ol4[i];
        const int4 mm = m4[i];
        float4 r;
        r.x = grad_of(a.x, b.x, (float)mm.x, o.x);
        r.y = grad_of(a.y, b.y, (float)mm.y, o.y);
        r.z = grad_of(a.z, b.z, (float)mm.z, o.z);
        r.w = grad_of(a.w, b.w, (float)mm.w, o.w);
        g4[i] = r;
      }
    } else {
#pragma unroll 4
      for (int64_t t = threadIdx.x; t < Tc; t += kGspoThreads)
        g_row[t] = grad_of(lp_row[t], ref_row[t], (float)m_row[t], old_row ? old_row[t] : 0.f);
    }
  }
  // ---- deterministic final reduction by the CTA that finishes the last sequence of the step (across range calls)
  __threadfence();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(p.ticket, 1u);
    s_last = (done == (unsigned int)(p.N - 1));
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float a = 0.f, b = 0.f;
    for (int64_t i = threadIdx.x; i < p.N; i += kGspoThreads) {
      a += __ldcg(p.seq_loss + i);
      b += __ldcg(p.seq_kl + i);
    }
    a = block_sum(a, s_red);
    b = block_sum(b, s_red);
    if (threadIdx.x == 0) {
      *p.loss = a / (float)p.N;                             // .mean() :706
      if (p.mean_kl) *p.mean_kl = b / (float)p.N;           // .mean() :737
      *p.ticket = 0u;                                       // re-arm for the next step
    }
  }
}

}  // namespace o3v

extern "C" int o3v_eos_mask(const int64_t* completion_ids, int64_t N, int64_t Tc, int64_t eos_id,
                            int64_t* eos_idx, int32_t* completion_mask, void* stream) {
  if (!completion_ids || !eos_idx || !completion_mask) return O3V_ERR_INVALID_ARG;
  if (N < 0 || Tc < 0 || Tc > 0x7fffffff) return O3V_ERR_INVALID_ARG;
  int rc = o3v::check_device();
  if (rc) return rc;
  if (N == 0) return O3V_OK;
  o3v::eos_mask_kernel<<<(unsigned)N, o3v::kEosThreads, 0, (cudaStream_t)stream>>>(
      completion_ids, Tc, eos_id, eos_idx, completion_mask);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}

extern "C" size_t o3v_gspo_workspace_bytes(int64_t N) {
  if (N < 0) return 0;
  return (size_t)(3 * N + 4 + N * o3v::kMaxSlices * 4) * sizeof(float);
}

extern "C" int o3v_gspo_fwd_bwd(const float* logp, const float* old_logp, const float* ref_logp,
                                const int32_t* mask, const float* rewards_per_func,
                                int64_t N, int64_t Tc, int64_t F, int64_t G,
                                int64_t seq_offset, int64_t n_seq,
                                float beta, float eps_low, float eps_high, int32_t gspo,
                                float* loss, float* mean_kl, float* advantages, float* reward_std,
                                int32_t* completion_len, float* grad_logp, float* per_token_kl,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (!logp || !ref_logp || !mask || !rewards_per_func || !loss || !workspace) return O3V_ERR_INVALID_ARG;
  if (N <= 0 || Tc <= 0 || F <= 0 || G <= 0 || (N % G) != 0) return O3V_ERR_INVALID_ARG;
  if (seq_offset < 0 || n_seq <= 0 || seq_offset + n_seq > N) return O3V_ERR_INVALID_ARG;
  if (workspace_bytes < o3v_gspo_workspace_bytes(N)) return O3V_ERR_WORKSPACE;
  int rc = o3v::check_device();
  if (rc) return rc;
  o3v::GspoParams p;
  p.logp = logp; p.old_logp = old_logp; p.ref = ref_logp; p.mask = mask; p.rpf = rewards_per_func;
  p.N = N; p.Tc = Tc; p.F = F; p.G = G; p.seq_offset = seq_offset;
  // CTAs per sequence: fill ~2 waves of SMs over the whole step, at least 512 tokens per CTA; a function
  // of (N, Tc) only so that every range call of a step uses the same slicing
  int64_t S = (2 * o3v::num_sms() + N - 1) / N;
  S = std::min<int64_t>(S, (Tc + 511) / 512);
  S = std::max<int64_t>(1, std::min<int64_t>(S, o3v::kMaxSlices));
  p.slice_len = (int32_t)((Tc + S - 1) / S);
  p.S = (int32_t)((Tc + p.slice_len - 1) / p.slice_len);
  p.beta = beta; p.eps_lo = eps_low; p.eps_hi = eps_high; p.gspo = gspo ? 1 : 0;
  p.loss = loss; p.mean_kl = mean_kl; p.adv = advantages; p.rstd = reward_std; p.clen = completion_len;
  p.grad = grad_logp; p.kl_out = per_token_kl;
  float* ws = (float*)workspace;
  p.seq_loss = ws; p.seq_kl = ws + N; p.ticket = (unsigned int*)(ws + 2 * N); p.seq_ticket = p.ticket + 4;
  p.partials = ws + 3 * N + 4;
  cudaStream_t st = (cudaStream_t)stream;
  // tickets: arrivals per sequence (the last slice CTA finishes the sequence) and finished sequences ACROSS the
  // range calls of one step (the CTA that finishes the last one reduces loss / mean_kl over all N in index order)
  if (seq_offset == 0) O3V_CUDA_TRY(cudaMemsetAsync(p.ticket, 0, (size_t)(4 + N) * sizeof(unsigned int), st));
  const dim3 grid((unsigned)p.S, (unsigned)n_seq);
  o3v::gspo_kernel<<<grid, o3v::kGspoThreads, 0, st>>>(p);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
