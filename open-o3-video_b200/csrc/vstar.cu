// K5: V-STAR scorer numerics (eval/test/eval_vstar.py:90-178 of the reference), fp64.
// One thread per (item, answer chain): the work per item is a few dozen IoUs, the point of the
// kernel is exactness (numpy's summation order, IEEE ops without FMA contraction) at batch scale.
#include "common.cuh"

namespace o3v {

constexpr int kVstarMaxF = 64;

__device__ __forceinline__ double vs_iou(const double* g, const double* p) {   // :112-133
  const double x1 = fmax(g[0], p[0]), y1 = fmax(g[1], p[1]);
  const double x2 = fmin(g[2], p[2]), y2 = fmin(g[3], p[3]);
  const double inter = __dmul_rn(fmax(0.0, __dsub_rn(x2, x1)), fmax(0.0, __dsub_rn(y2, y1)));
  const double ga = __dmul_rn(__dsub_rn(g[2], g[0]), __dsub_rn(g[3], g[1]));
  const double pa = __dmul_rn(__dsub_rn(p[2], p[0]), __dsub_rn(p[3], p[1]));
  const double uni = __dsub_rn(__dadd_rn(ga, pa), inter);
  return uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
}

// numpy's pairwise summation for n <= 128 (np.mean of a float64 array): n < 8 sequential from 0.0,
// else 8 interleaved accumulators combined as a tree, then the remainder sequentially.
__device__ double numpy_sum(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
    return r;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

__global__ void vstar_kernel(const o3v_vstar_soa s, double* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= s.I * 2) return;
  const int64_t i = tid >> 1;
  const int c = (int)(tid & 1);
  double* o = out + i * 14 + c * 7;
  // ---- temporal IoU (:90-109)
  double tiou = 0.0;
  if (s.t_valid[i * 2 + c]) {
    const double gs = s.gt_seg[i * 2], ge = s.gt_seg[i * 2 + 1];
    const double ps = s.pred_seg[(i * 2 + c) * 2], pe = s.pred_seg[(i * 2 + c) * 2 + 1];
    const double inter = fmax(0.0, __dsub_rn(fmin(ge, pe), fmax(gs, ps)));
    const double uni = __dsub_rn(fmax(ge, pe), fmin(gs, ps));
    tiou = uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
  }
  o[0] = tiou;
  // ---- spatial metrics (:148-178)
  double miou = 0.0, aps[5] = {0, 0, 0, 0, 0};
  const int nf = s.n_frames[i];
  if (s.sp_valid[i * 2 + c] && nf > 0) {
    double ious[kVstarMaxF];
    int hits[5] = {0, 0, 0, 0, 0};
    const double thr[5] = {0.1, 0.3, 0.5, 0.7, 0.9};
    for (int f = 0; f < nf; ++f) {
      const int64_t cf = (i * 2 + c) * s.F + f;
      const int nb = s.n_pb[cf];
      const unsigned valid = s.pb_valid[cf];
      const double* g = s.gt_box + (i * s.F + f) * 4;
      double best = 0.0;
      for (int b = 0; b < nb; ++b) {                       // max([...]) over the frame's predictions (:143)
        double v = 0.0;
        if ((valid >> b) & 1u) v = vs_iou(g, s.pb + (cf * s.Pb + b) * 4);
        best = (b == 0) ? v : fmax(best, v);
      }
      ious[f] = best;
      for (int k = 0; k < 5; ++k) hits[k] += (best >= thr[k]) ? 1 : 0;
    }
    miou = __ddiv_rn(numpy_sum(ious, nf), (double)nf);     // np.mean (:167)
    for (int k = 0; k < 5; ++k) aps[k] = __ddiv_rn((double)hits[k], (double)nf);   // :170-173
  }
  o[1] = miou;
  for (int k = 0; k < 5; ++k) o[2 + k] = aps[k];
}

}  // namespace o3v

extern "C" int o3v_vstar_scores(const o3v_vstar_soa* soa, double* out, void* stream) {
  if (!soa || !out) return O3V_ERR_INVALID_ARG;
  const o3v_vstar_soa& s = *soa;
  if (s.I < 0 || s.F < 0 || s.Pb < 0) return O3V_ERR_INVALID_ARG;
  if (s.F > o3v::kVstarMaxF || s.Pb > 32) return O3V_ERR_SHAPE;
  if (!s.t_valid || !s.gt_seg || !s.pred_seg || !s.sp_valid || !s.n_frames || !s.gt_box || !s.n_pb ||
      !s.pb_valid || !s.pb)
    return O3V_ERR_INVALID_ARG;
  int rc = o3v::check_device();
  if (rc) return rc;
  if (s.I == 0) return O3V_OK;
  const int threads = 128;
  const unsigned grid = (unsigned)((s.I * 2 + threads - 1) / threads);
  o3v::vstar_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(s, out);
  O3V_LAUNCH_CHECK();
  return O3V_OK;
}
