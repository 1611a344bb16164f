// TEST FIXTURE, not a product path: compiles csrc/scan_core.cuh (the K6 scanner, host + device
// code) with g++ so that the CPU test-suite can fuzz the scanner's logic against Python's
// re / json / float on the build box, where there is no GPU.  The package never loads this
// library; the product path is the CUDA kernel in csrc/parse.cu (lib/libo3v.so).
#include <stdint.h>
#include <string.h>

#include "../../open-o3-video_b200/csrc/scan_core.cuh"

using namespace o3v::scan;

extern "C" {

// float(str.strip()) on bytes [0, n): 1 = ok
int scan_host_python_float(const uint8_t* s, int64_t n, int strip_mode, double* out) {
  int64_t a = 0, b = n;
  strip_space(s, &a, &b, strip_mode != 0);
  return python_float(s, a, b, out) ? 1 : 0;
}

double scan_host_dec(const uint8_t* s, int64_t n, int64_t exp10) { return dec_to_double(s, 0, n, exp10); }
double scan_host_big(const uint8_t* s, int64_t n, int64_t exp10) { return big_to_double(s, 0, n, exp10); }

int scan_host_json_box(const uint8_t* s, int64_t n, int* n_elems, int* numeric, double* out4) {
  bool num = false;
  int ne = 0;
  double v[4] = {0, 0, 0, 0};
  const int rc = json_box(s, 0, n, &ne, &num, v);
  *n_elems = ne;
  *numeric = num ? 1 : 0;
  memcpy(out4, v, sizeof(v));
  return rc;
}

// same argument block as o3v_parse_args, host pointers
struct Args {
  int64_t R, G;
  int32_t P, C, Bc, Tb;
  const uint8_t* text; const int64_t* offsets; const int32_t* task;
  int32_t* flags; double* ans_seg; double* ans_box; int32_t* n_times; double* think_times; int32_t* n_claims;
  double* claim_t; int32_t* claim_nbox; uint32_t* claim_valid; double* claim_box; int32_t* n_tboxes;
  uint32_t* tbox_valid; double* think_box; int32_t* overflow;
};

int scan_host_parse(const Args* ap) {
  const Args& a = *ap;
  for (int i = 0; i < 4; ++i) a.overflow[i] = 0;
  for (int64_t r = 0; r < a.R; ++r) {
    Caps cap{a.P, a.C, a.Bc, a.Tb};
    RolloutOut o;
    o.flags = a.flags + r; o.ans_seg = a.ans_seg + r * 2; o.ans_box = a.ans_box + r * 4;
    o.n_times = a.n_times + r; o.think_times = a.think_times + r * a.P; o.n_claims = a.n_claims + r;
    o.claim_t = a.claim_t + r * a.C; o.claim_nbox = a.claim_nbox + r * a.C; o.claim_valid = a.claim_valid + r * a.C;
    o.claim_box = a.claim_box + r * (int64_t)a.C * a.Bc * 4; o.n_tboxes = a.n_tboxes + r;
    o.tbox_valid = a.tbox_valid + r; o.think_box = a.think_box + r * (int64_t)a.Tb * 4;
    Scratch sc;
    scan_rollout(a.text, a.offsets[a.R], a.offsets[r], a.offsets[r + 1], a.task[r / a.G], cap, o, &sc, nullptr);
    for (int item = 0; item < items_per_rollout(cap); ++item)
      if (item_active(item, cap, o, &sc)) convert_item(a.text, item, cap, o, &sc, 0xffffffffu);
    int over[4];
    finish_rollout(cap, o, &sc, over);
    for (int i = 0; i < 4; ++i)
      if (over[i] > a.overflow[i]) a.overflow[i] = over[i];
  }
  return 0;
}

}  // extern "C"
