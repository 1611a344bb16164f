// K6 core: byte-level scanner that turns completion text into the rollout side of
// o3v_rewards_soa (the step before K4).
//
// Replaces, bit for bit, the regex / json / float() extraction of the reference's
// src/r1-v/src/open_r1/reward_func.py:
//   :91-93, :189 extract_answer      :119 / :149  answer "<t>a</t>s to <t>b</t>s"
//   :211-223     answer <box>        :308-335     parse_temporal_spatial_reasoning_process
//   :394, :437, :481-482 think/answer spans      :405-412, :447-449 think "<t>x</t>s"
//   :492-511     visual-QA think boxes
//
// The reference's patterns are lazy-quantifier regexes over literal tags.  Each of them reduces to
// an ordered sequence of "first occurrence of literal X at or after p" searches (derivation in
// DESIGN.md section 8); numbers go through an exactly rounded decimal -> binary64 conversion (what
// Python's float() and json.loads produce) and box payloads through a JSON recogniser that accepts
// what CPython's C scanner accepts.
//
// Everything here is `O3V_HD` (host + device).  The device build (parse.cu) runs three phases so
// that no phase leaves SIMT lanes idle behind divergent nested loops:
//   A  scan_rollout   one WARP per rollout: warp-cooperative literal searches (16 bytes per lane,
//                     512 bytes per step) resolve the spans and the regex match chains and record
//                     CANDIDATES as byte ranges (packed into the output rows themselves);
//   B  convert_item   one THREAD per candidate: the sequential number / JSON routines, every lane
//                     of a warp running the same routine on its own range;
//   C  finish_rollout one thread per rollout: drops rejected candidates, compacts, final counts.
// The host build of this same header (tests/hostbuild/) runs A, B, C in plain loops so that the
// CPU tests can fuzz the logic against Python; the product path is the CUDA build.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define O3V_HD __host__ __device__ __forceinline__
#define O3V_HD_NOINLINE __host__ __device__ __noinline__
#define O3V_TABLE static __device__ const
#define O3V_CONSTEXPR __host__ __device__ constexpr
#else
#define O3V_CONSTEXPR constexpr
#define O3V_HD inline
#define O3V_HD_NOINLINE inline
#endif

// A data-dependent loop run by the lanes of a warp on DIFFERENT data (phase B).  `lanes` is the set of lanes
// that run the loop (all of them must reach it); the vote at the top makes them start every iteration together
// instead of drifting apart until the loop's exit (finished lanes idle until the last one is).  The body must
// leave only through `continue` (no break / return).  O3V_LOCKSTEP_WHO(p) = the running lanes for which p holds;
// every running lane must evaluate it (no `continue` before it).
#if defined(__CUDA_ARCH__)
#define O3V_LOCKSTEP_BEGIN(lanes, cond)                                        \
  {                                                                            \
    const unsigned lockstep_all__ = (lanes);                                   \
    for (;;) {                                                                 \
      const bool lockstep_go__ = (cond);                                       \
      const unsigned lockstep_run__ = __ballot_sync(lockstep_all__, lockstep_go__); \
      if (lockstep_run__ == 0) break;                                          \
      if (lockstep_go__) {
#define O3V_LOCKSTEP_WHO(pred) __ballot_sync(lockstep_run__, (pred))
#define O3V_LOCKSTEP_END \
      }                  \
    }                    \
  }
#define O3V_SYNC_LANES(lanes) __syncwarp(lanes)
#else
#define O3V_LOCKSTEP_BEGIN(lanes, cond) \
  {                                     \
    (void)(lanes);                      \
    for (;;) {                          \
      if (!(cond)) break;               \
      {
#define O3V_LOCKSTEP_WHO(pred) ((pred) ? 1u : 0u)
#define O3V_LOCKSTEP_END \
      }                  \
    }                    \
  }
#define O3V_SYNC_LANES(lanes) (void)(lanes)
#endif

namespace o3v {
namespace scan {

// ------------------------------------------------------------------------------------------
// Tables.  Unicode 15.0 decimal digits (category Nd; what Python 3.12's `\d` and float() accept):
// first code point of every run of ten, value = (cp - start) % 10 (generated and checked by
// tools/gen_unicode_tables.py).  The mathematical digits U+1D7CE..1D7FF are five runs of ten.
// ------------------------------------------------------------------------------------------
#define O3V_ND_STARTS                                                                                        \
  {0x0660, 0x06F0, 0x07C0, 0x0966, 0x09E6, 0x0A66, 0x0AE6, 0x0B66, 0x0BE6, 0x0C66, 0x0CE6, 0x0D66, 0x0DE6,   \
   0x0E50, 0x0ED0, 0x0F20, 0x1040, 0x1090, 0x17E0, 0x1810, 0x1946, 0x19D0, 0x1A80, 0x1A90, 0x1B50, 0x1BB0,   \
   0x1C40, 0x1C50, 0xA620, 0xA8D0, 0xA900, 0xA9D0, 0xA9F0, 0xAA50, 0xABF0, 0xFF10, 0x104A0, 0x10D30,         \
   0x11066, 0x110F0, 0x11136, 0x111D0, 0x112F0, 0x11450, 0x114D0, 0x11650, 0x116C0, 0x11730, 0x118E0,        \
   0x11950, 0x11C50, 0x11D50, 0x11DA0, 0x11F50, 0x16A60, 0x16AC0, 0x16B50, 0x1D7CE, 0x1D7D8, 0x1D7E2,        \
   0x1D7EC, 0x1D7F6, 0x1E140, 0x1E2F0, 0x1E4F0, 0x1E950, 0x1FBF0}
constexpr int kNdRuns = 67;
#define O3V_POW10                                                                                            \
  {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18,   \
   1e19, 1e20, 1e21, 1e22}

static const uint32_t kNdStartsHost[kNdRuns] = O3V_ND_STARTS;
static const double kPow10Host[23] = O3V_POW10;
#if defined(__CUDACC__)
O3V_TABLE uint32_t kNdStartsDev[kNdRuns] = O3V_ND_STARTS;
O3V_TABLE double kPow10Dev[23] = O3V_POW10;
#endif

O3V_HD const uint32_t* nd_starts() {
#if defined(__CUDA_ARCH__)
  return kNdStartsDev;
#else
  return kNdStartsHost;
#endif
}
O3V_HD double pow10_exact(int e) {
#if defined(__CUDA_ARCH__)
  return kPow10Dev[e];
#else
  return kPow10Host[e];
#endif
}

// ------------------------------------------------------------------------------------------
// Literals, packed at compile time (no memory traffic for the needle).
// ------------------------------------------------------------------------------------------
struct Lit {
  uint64_t lo, hi;
  int n;
};
template <int N>
O3V_CONSTEXPR Lit make_lit(const char (&s)[N]) {
  Lit l{0, 0, N - 1};
  for (int i = 0; i < N - 1; ++i) {
    if (i < 8) l.lo |= (uint64_t)(uint8_t)s[i] << (8 * i);
    else l.hi |= (uint64_t)(uint8_t)s[i] << (8 * (i - 8));
  }
  return l;
}
O3V_HD uint8_t lit_byte(const Lit& l, int i) { return (uint8_t)((i < 8 ? l.lo >> (8 * i) : l.hi >> (8 * (i - 8))) & 0xff); }

// t[p .. p+n) == literal, with p + n <= end
O3V_HD bool lit_at(const uint8_t* t, int64_t p, int64_t end, const Lit& l) {
  if (p < 0 || p + l.n > end) return false;
  for (int i = 0; i < l.n; ++i)
    if (t[p + i] != lit_byte(l, i)) return false;
  return true;
}

// The literals the match chains search for.
enum LitId { kLThinkO = 0, kLThinkC, kLAnsO, kLAnsC, kLT, kLTEnd, kLObj, kLObjBox, kLBoxAt, kLBoxOpen, kLBoxClose,
             kLNewline, kNumLits };
template <int ID>
O3V_CONSTEXPR Lit lit_of() {
  return ID == kLThinkO ? make_lit("<think>") : ID == kLThinkC ? make_lit("</think>") :
         ID == kLAnsO ? make_lit("<answer>") : ID == kLAnsC ? make_lit("</answer>") :
         ID == kLT ? make_lit("<t>") : ID == kLTEnd ? make_lit("</t>s") : ID == kLObj ? make_lit("<obj>") :
         ID == kLObjBox ? make_lit("</obj><box>[") : ID == kLBoxAt ? make_lit("]</box>at<t>") :
         ID == kLBoxOpen ? make_lit("<box>[") : ID == kLBoxClose ? make_lit("]</box>") : make_lit("\n");
}

// find<ID>(from, end): first p in [from, end - n] with t[p .. p+n) == literal ID, else -1.
//
// Host: a plain scan.  Device (phase A, called by all 32 lanes of a warp with identical
// arguments): the warp keeps one 512-byte block of the text classified -- per lane, a 16-bit
// match mask per literal for its 16 positions -- so that the chains' many searches cost a mask
// select and one warp reduction each, and every block is loaded and classified once per sweep.
// `t` is the 16-byte aligned base of the whole text buffer (positions are absolute), `total` its
// length; the buffer is readable up to the next multiple of 16.
struct Finder {
  const uint8_t* t;
  int64_t total;
  static constexpr int kSmemBlocks = 8;                       // blocks of a rollout whose masks stay in smem
  static constexpr int kMaskWords = kSmemBlocks * (kNumLits / 2) * 32;
  static constexpr int kSmemWords = kMaskWords + 32;           // per warp: mask cache + a 32-entry position list
#if defined(__CUDA_ARCH__)
  int64_t base;              // start of the block classified in registers (multiple of 512), -1 = none
  uint32_t mask[kNumLits / 2];   // two 16-bit match masks per register
  uint32_t* sm;              // this warp's mask cache: [block][word][lane]
  int64_t first;             // block 0 of the cache = the rollout's first block
  uint32_t cached;           // bit b: block b of the rollout is in `sm`

  __device__ __forceinline__ void init(const uint8_t* text, int64_t n, int64_t rollout_beg, uint32_t* warp_smem) {
    t = text; total = n; base = -1; sm = warp_smem; first = rollout_beg & ~(int64_t)511; cached = 0;
  }

  template <int ID>
  __device__ __forceinline__ static bool win_is(uint32_t w0, uint32_t w1, uint32_t w2) {
    constexpr Lit l = lit_of<ID>();
    constexpr uint32_t a0 = (uint32_t)l.lo, a1 = (uint32_t)(l.lo >> 32), a2 = (uint32_t)l.hi;
    constexpr uint32_t m0 = l.n >= 4 ? 0xffffffffu : (1u << (8 * l.n)) - 1u;
    constexpr uint32_t m1 = l.n >= 8 ? 0xffffffffu : l.n <= 4 ? 0u : (1u << (8 * (l.n - 4))) - 1u;
    constexpr uint32_t m2 = l.n >= 12 ? 0xffffffffu : l.n <= 8 ? 0u : (1u << (8 * (l.n - 8))) - 1u;
    return (((w0 ^ a0) & m0) | ((w1 ^ a1) & m1) | ((w2 ^ a2) & m2)) == 0;
  }
  __device__ __forceinline__ static uint32_t eq16(const uint4& v, uint8_t c) {   // bit i: byte i == c
    const uint32_t c4 = (uint32_t)c * 0x01010101u;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      m |= ((((__vcmpeq4(w[i], c4) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << (4 * i);
    return m;
  }
  __device__ __forceinline__ void set(int id, int j) { mask[id >> 1] |= 1u << (j + 16 * (id & 1)); }

  // load and classify the block at `blk`
  __device__ __forceinline__ void load_block(int64_t blk) {
    const int lane = threadIdx.x & 31;
    const int64_t p0 = blk + lane * 16;
    const int64_t limit = (total + 15) & ~(int64_t)15;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (p0 < limit) v = *reinterpret_cast<const uint4*>(t + p0);
    // the 16 bytes after mine: the next lane's vector (the last lane fetches its own)
    uint4 nx;
    nx.x = __shfl_down_sync(0xffffffffu, v.x, 1); nx.y = __shfl_down_sync(0xffffffffu, v.y, 1);
    nx.z = __shfl_down_sync(0xffffffffu, v.z, 1); nx.w = __shfl_down_sync(0xffffffffu, v.w, 1);
    if (lane == 31) {
      nx = make_uint4(0, 0, 0, 0);
      if (p0 + 16 < limit) nx = *reinterpret_cast<const uint4*>(t + p0 + 16);
    }
#pragma unroll
    for (int i = 0; i < kNumLits / 2; ++i) mask[i] = 0;
    const uint32_t w[8] = {v.x, v.y, v.z, v.w, nx.x, nx.y, nx.z, nx.w};
    uint32_t cand = eq16(v, '<') | eq16(v, ']');
    mask[kLNewline >> 1] = eq16(v, '\n') << (16 * (kLNewline & 1));
    while (cand) {
      const int j = __ffs(cand) - 1;
      cand &= cand - 1;
      const int sh = (j & 3) * 8;
      uint32_t w0, w1, w2;                 // the 12 bytes starting at position j
      switch (j >> 2) {
        case 0: w0 = __funnelshift_r(w[0], w[1], sh); w1 = __funnelshift_r(w[1], w[2], sh); w2 = __funnelshift_r(w[2], w[3], sh); break;
        case 1: w0 = __funnelshift_r(w[1], w[2], sh); w1 = __funnelshift_r(w[2], w[3], sh); w2 = __funnelshift_r(w[3], w[4], sh); break;
        case 2: w0 = __funnelshift_r(w[2], w[3], sh); w1 = __funnelshift_r(w[3], w[4], sh); w2 = __funnelshift_r(w[4], w[5], sh); break;
        default: w0 = __funnelshift_r(w[3], w[4], sh); w1 = __funnelshift_r(w[4], w[5], sh); w2 = __funnelshift_r(w[5], w[6], sh); break;
      }
      if ((w0 & 0xff) == '<') {
        switch ((w0 >> 8) & 0xff) {                          // second byte picks the few tags worth testing
          case '/':
            if (win_is<kLTEnd>(w0, w1, w2)) set(kLTEnd, j);
            else if (win_is<kLObjBox>(w0, w1, w2)) set(kLObjBox, j);
            else if (win_is<kLThinkC>(w0, w1, w2)) set(kLThinkC, j);
            else if (win_is<kLAnsC>(w0, w1, w2)) set(kLAnsC, j);
            break;
          case 't':
            if (win_is<kLT>(w0, w1, w2)) set(kLT, j);
            else if (win_is<kLThinkO>(w0, w1, w2)) set(kLThinkO, j);
            break;
          case 'o': if (win_is<kLObj>(w0, w1, w2)) set(kLObj, j); break;
          case 'b': if (win_is<kLBoxOpen>(w0, w1, w2)) set(kLBoxOpen, j); break;
          case 'a': if (win_is<kLAnsO>(w0, w1, w2)) set(kLAnsO, j); break;
          default: break;
        }
      } else if (win_is<kLBoxClose>(w0, w1, w2)) {
        set(kLBoxClose, j);
        if (win_is<kLBoxAt>(w0, w1, w2)) set(kLBoxAt, j);
      }
    }
    base = blk;
  }

  // my 16-bit match mask of literal ID in block `blk`, restricted to start positions in [from, last]
  template <int ID>
  __device__ __forceinline__ uint32_t block_mask(int64_t blk, int64_t from, int64_t last) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (blk - first) >> 9;
    uint32_t word;
    if (b >= 0 && b < kSmemBlocks) {                         // classified once per rollout, then served from smem
      uint32_t* row = sm + (int)b * (kNumLits / 2) * 32 + lane;
      if (!((cached >> (int)b) & 1u)) {
        load_block(blk);
#pragma unroll
        for (int i = 0; i < kNumLits / 2; ++i) row[i * 32] = mask[i];
        cached |= 1u << (int)b;
        word = mask[ID >> 1];
      } else {
        word = row[(ID >> 1) * 32];
      }
    } else {
      if (blk != base) load_block(blk);
      word = mask[ID >> 1];
    }
    const int64_t p0 = blk + lane * 16;
    uint32_t m = (word >> (16 * (ID & 1))) & 0xffffu;
    const int64_t lo = from - p0, hi = last - p0;            // keep positions lo..hi of my 16
    if (lo >= 16 || hi < 0) return 0;
    if (lo > 0) m &= 0xffffu << (int)lo;
    if (hi < 15) m &= (2u << (int)hi) - 1u;
    return m;
  }

  template <int ID>
  __device__ __forceinline__ int64_t find(int64_t from, int64_t end) {
    constexpr int n = lit_of<ID>().n;
    const int64_t last = end - n;
    if (from < 0 || from > last) return -1;
    const int lane = threadIdx.x & 31;
    for (;;) {
      const int64_t blk = from & ~(int64_t)511;
      const uint32_t m = block_mask<ID>(blk, from, last);
      uint32_t best = m ? (uint32_t)(lane * 16 + __ffs(m) - 1) : 0xffffffffu;
      best = __reduce_min_sync(0xffffffffu, best);
      if (best != 0xffffffffu) return blk + best;
      from = blk + 512;
      if (from > last) return -1;
    }
  }

  // Per-LANE search in the shared-memory masks (every block of [from, end) must already be cached and
  // published with __syncwarp): first start position >= from of literal ID, else -1.
  template <int ID>
  __device__ __forceinline__ int64_t next_lane(int64_t from, int64_t end) const {
    const int64_t last = end - lit_of<ID>().n;
    if (from > last) return -1;
    uint32_t rel = (uint32_t)(from - first);                  // < kSmemBlocks * 512
    const uint32_t rel_last = (uint32_t)(last - first);
    for (;;) {
      const uint32_t word = sm[(rel >> 9) * (kNumLits / 2) * 32 + (ID >> 1) * 32 + ((rel >> 4) & 31)];
      const uint32_t m = ((word >> (16 * (ID & 1))) & 0xffffu) & (0xffffu << (rel & 15));
      if (m) {
        const uint32_t p = (rel & ~15u) + (uint32_t)(__ffs(m) - 1);
        return p <= rel_last ? first + p : -1;
      }
      rel = (rel & ~15u) + 16;
      if (rel > rel_last) return -1;
    }
  }
#else
  void init(const uint8_t* text, int64_t n, int64_t, uint32_t*) { t = text; total = n; }
  template <int ID>
  int64_t find(int64_t from, int64_t end) const {
    constexpr Lit l = lit_of<ID>();
    const int64_t last = end - l.n;
    if (from < 0) return -1;
    const uint8_t c0 = lit_byte(l, 0);
    for (int64_t p = from; p <= last; ++p)
      if (t[p] == c0 && lit_at(t, p, end, l)) return p;
    return -1;
  }
#endif
};

// ------------------------------------------------------------------------------------------
// UTF-8 helpers.  The text is Python str encoded as UTF-8 (well formed); malformed bytes are
// treated as single opaque characters.
// ------------------------------------------------------------------------------------------
O3V_HD uint32_t decode_utf8(const uint8_t* t, int64_t p, int64_t end, int* len) {
  const uint32_t b0 = t[p];
  if (b0 < 0x80) { *len = 1; return b0; }
  if (b0 >= 0xC2 && b0 <= 0xDF && p + 1 < end) { *len = 2; return ((b0 & 0x1F) << 6) | (t[p + 1] & 0x3F); }
  if (b0 >= 0xE0 && b0 <= 0xEF && p + 2 < end) {
    *len = 3;
    return ((b0 & 0x0F) << 12) | ((uint32_t)(t[p + 1] & 0x3F) << 6) | (t[p + 2] & 0x3F);
  }
  if (b0 >= 0xF0 && b0 <= 0xF4 && p + 3 < end) {
    *len = 4;
    return ((b0 & 0x07) << 18) | ((uint32_t)(t[p + 1] & 0x3F) << 12) | ((uint32_t)(t[p + 2] & 0x3F) << 6) |
           (t[p + 3] & 0x3F);
  }
  *len = 1;
  return 0xFFFFFFFFu;
}

// Unicode decimal digit at p: value 0..9 and its byte length, else -1.
O3V_HD int digit_at(const uint8_t* t, int64_t p, int64_t end, int* len) {
  if (p >= end) return -1;
  const uint8_t b = t[p];
  if (b < 0x80) {
    *len = 1;
    return (b >= '0' && b <= '9') ? (int)(b - '0') : -1;
  }
  const uint32_t cp = decode_utf8(t, p, end, len);
  const uint32_t* starts = nd_starts();
  for (int i = 0; i < kNdRuns; ++i)
    if (cp - starts[i] < 10u) return (int)(cp - starts[i]);
  return -1;
}

// str.isspace() set (what re's \s and str.strip() treat as whitespace).  float() itself strips the
// same set minus the separators U+001C..U+001F (`strict` = false).
O3V_HD bool is_space_cp(uint32_t cp, bool strip_mode) {
  if (cp >= 0x1C && cp <= 0x1F) return strip_mode;
  return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
         (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F ||
         cp == 0x3000;
}
// [s, e) -> stripped of leading / trailing Unicode whitespace (strip_mode: str.strip(), else float()'s own)
O3V_HD void strip_space(const uint8_t* t, int64_t* s, int64_t* e, bool strip_mode) {
  int64_t a = *s, b = *e;
  while (a < b) {
    int len;
    const uint32_t cp = decode_utf8(t, a, b, &len);
    if (!is_space_cp(cp, strip_mode)) break;
    a += len;
  }
  while (b > a) {
    int64_t q = b - 1;                       // start of the last character
    while (q > a && (t[q] & 0xC0) == 0x80 && b - q < 4) --q;
    int len;
    const uint32_t cp = decode_utf8(t, q, b, &len);
    if (q + len != b || !is_space_cp(cp, strip_mode)) break;
    b = q;
  }
  *s = a;
  *e = b;
}

// ------------------------------------------------------------------------------------------
// Exact decimal -> binary64.
//
// A number is handed over as a mantissa range of the text (digits of any Nd script, at most one
// '.', '_' separators; the caller has validated the grammar) plus an explicit power of ten.
// Fast path (Clinger): <= 19 significant digits, integer < 2^53, |exponent| <= 22 -> one exact
// multiply or divide.  Everything else goes through a big decimal (800 digits + sticky flag,
// enough for any binary64 rounding decision) that is scaled by powers of two until it lies in
// [1/2, 1), then 53 bits are extracted and rounded half-even.
// ------------------------------------------------------------------------------------------
constexpr int kBigDigits = 800;
struct BigDec {
  uint8_t d[kBigDigits];
  int nd;       // digits in use
  int64_t dp;   // value = 0.d0 d1 ... x 10^dp
  bool sticky;  // non-zero digits were dropped below d[nd-1]
};

O3V_HD void big_trim(BigDec& b) {
  while (b.nd > 0 && b.d[b.nd - 1] == 0) --b.nd;
  if (b.nd == 0) b.dp = 0;
}

// b /= 2^k, 1 <= k <= 60
O3V_HD void big_shr(BigDec& b, int k) {
  int r = 0, w = 0;
  uint64_t n = 0;
  while ((n >> k) == 0) {
    if (r >= b.nd) {
      if (n == 0) { b.nd = 0; b.dp = 0; return; }
      while ((n >> k) == 0) { n *= 10; ++r; }
      break;
    }
    n = n * 10 + b.d[r++];
  }
  b.dp -= r - 1;
  const uint64_t mask = ((uint64_t)1 << k) - 1;
  for (; r < b.nd; ++r) {
    const uint64_t dig = n >> k;
    n &= mask;
    b.d[w++] = (uint8_t)dig;
    n = n * 10 + b.d[r];
  }
  while (n > 0) {
    const uint64_t dig = n >> k;
    n &= mask;
    if (w < kBigDigits) b.d[w++] = (uint8_t)dig;
    else if (dig > 0) b.sticky = true;
    n *= 10;
  }
  b.nd = w;
  big_trim(b);
}

// b *= 2^k, 1 <= k <= 60
O3V_HD void big_shl(BigDec& b, int k) {
  if (b.nd == 0) return;
  const int delta = (k * 30103) / 100000 + 1;   // >= number of new leading digits
  int w = b.nd + delta - 1;
  uint64_t n = 0;
  for (int r = b.nd - 1; r >= 0; --r, --w) {
    n += (uint64_t)b.d[r] << k;
    const uint64_t q = n / 10, rem = n - 10 * q;
    if (w < kBigDigits) b.d[w] = (uint8_t)rem;
    else if (rem) b.sticky = true;
    n = q;
  }
  for (; n > 0; --w) {
    const uint64_t q = n / 10, rem = n - 10 * q;
    if (w < kBigDigits) b.d[w] = (uint8_t)rem;
    else if (rem) b.sticky = true;
    n = q;
  }
  const int lead = w + 1;                       // unused positions in front
  int nd = b.nd + delta - lead;
  if (nd > kBigDigits - lead) nd = kBigDigits - lead;
  if (lead > 0)
    for (int i = 0; i < nd; ++i) b.d[i] = b.d[i + lead];
  b.nd = nd;
  b.dp += delta - lead;
  big_trim(b);
}

O3V_HD void big_shift(BigDec& b, int k) {   // k > 0: multiply by 2^k, k < 0: divide
  while (k > 60) { big_shl(b, 60); k -= 60; }
  while (k < -60) { big_shr(b, 60); k += 60; }
  if (k > 0) big_shl(b, k);
  else if (k < 0) big_shr(b, -k);
}

// integer part of b (dp <= 18) rounded half-even on the dropped digits (+ sticky)
O3V_HD uint64_t big_rounded_integer(const BigDec& b) {
  uint64_t n = 0;
  int i = 0;
  for (; i < b.dp && i < b.nd; ++i) n = n * 10 + b.d[i];
  for (; i < b.dp; ++i) n *= 10;
  bool up = false;
  if (b.dp >= 0 && b.dp < b.nd) {
    const int idx = (int)b.dp;
    if (b.d[idx] == 5 && idx + 1 == b.nd) up = b.sticky || (idx > 0 && (b.d[idx - 1] & 1));   // exactly half: even
    else up = b.d[idx] >= 5;
  }
  return n + (up ? 1 : 0);
}

// In-slot markers of box validity (include/o3v.h: O3V_INVALID_BOX_BITS).  Quiet NaNs with a payload that neither
// float() / json.loads (canonical 0x7FF8000000000000) nor a packed candidate range (text offsets < 2^33) can produce.
constexpr uint64_t kInvalidBoxBits = 0x7FF8B0B0DEADBEEFull;   // a JSON list that is not 4 numbers
constexpr uint64_t kDroppedBoxBits = 0x7FF8B0B0DEADD00Dull;   // not JSON at all (internal to K6: never leaves phase C)

O3V_HD uint64_t double_to_bits(double v) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(v);
#else
  union { uint64_t u; double d; } c;
  c.d = v;
  return c.u;
#endif
}

O3V_HD double bits_to_double(uint64_t bits) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)bits);
#else
  union { uint64_t u; double d; } c;
  c.u = bits;
  return c.d;
#endif
}

// Slow path: rebuild the digits from the text.  Kept out of line: it needs an 800-byte frame.
O3V_HD_NOINLINE double big_to_double(const uint8_t* t, int64_t ms, int64_t me, int64_t exp10) {
  BigDec b;
  b.nd = 0; b.dp = 0; b.sticky = false;
  int64_t int_digits = 0, lead_zeros = 0;
  bool seen_point = false, seen_nonzero = false;
  for (int64_t p = ms; p < me;) {
    const uint8_t c = t[p];
    if (c == '.') { seen_point = true; ++p; continue; }
    if (c == '_') { ++p; continue; }
    int len = 1;
    const int v = digit_at(t, p, me, &len);
    p += len;
    if (v < 0) continue;                     // not reachable for validated input
    if (!seen_point) ++int_digits;
    if (!seen_nonzero) {
      if (v == 0) { ++lead_zeros; continue; }
      seen_nonzero = true;
    }
    if (b.nd < kBigDigits) b.d[b.nd++] = (uint8_t)v;
    else if (v) b.sticky = true;
  }
  if (b.nd == 0) return 0.0;
  b.dp = int_digits - lead_zeros + exp10;
  big_trim(b);
  if (b.dp > 310) return bits_to_double(0x7FF0000000000000ull);
  if (b.dp < -330) return 0.0;
  const int tab[9] = {0, 3, 6, 9, 13, 16, 19, 23, 26};
  int exp2 = 0;
  while (b.dp > 0) {
    const int n = b.dp >= 9 ? 27 : tab[b.dp];
    big_shr(b, n);
    exp2 += n;
  }
  while (b.dp < 0 || (b.dp == 0 && b.d[0] < 5)) {
    const int n = -b.dp >= 9 ? 27 : (b.dp == 0 ? 1 : tab[-b.dp]);
    big_shl(b, n);
    exp2 -= n;
  }
  exp2 -= 1;                                  // value = [1/2, 1) x 2^(exp2+1) = [1, 2) x 2^exp2
  if (exp2 < -1022) {                         // subnormal: fewer mantissa bits
    const int n = -1022 - exp2;
    big_shift(b, -n);
    exp2 += n;
  }
  if (exp2 > 1023) return bits_to_double(0x7FF0000000000000ull);
  big_shift(b, 53);
  uint64_t mant = big_rounded_integer(b);
  if (mant == ((uint64_t)2 << 52)) {
    mant >>= 1;
    ++exp2;
    if (exp2 > 1023) return bits_to_double(0x7FF0000000000000ull);
  }
  uint64_t biased = (uint64_t)(exp2 + 1023);
  if ((mant & ((uint64_t)1 << 52)) == 0) biased = 0;   // subnormal (or zero)
  return bits_to_double((mant & (((uint64_t)1 << 52) - 1)) | (biased << 52));
}

// Digit accumulator of the fast path: the scanners feed it while they recognise a number, so the text
// is walked once; only the slow path re-reads the mantissa range.
struct DecAcc {
  uint64_t w;        // first 19 significant digits
  int nsig;
  int64_t frac, dropped;   // digits after the point / beyond the 19th
  bool exact;        // no non-zero digit was dropped
};
O3V_HD void acc_init(DecAcc& a) { a.w = 0; a.nsig = 0; a.frac = 0; a.dropped = 0; a.exact = true; }
O3V_HD void acc_digit(DecAcc& a, int v, bool after_point) {
  if (after_point) ++a.frac;
  if (a.nsig == 0 && v == 0) return;          // leading zero
  if (a.nsig < 19) { a.w = a.w * 10 + (uint64_t)v; ++a.nsig; }
  else { ++a.dropped; if (v) a.exact = false; }
}
// value of the accumulated digits x 10^exp10 (unsigned), correctly rounded; [ms, me) is the mantissa text
O3V_HD double acc_value(const DecAcc& a, const uint8_t* t, int64_t ms, int64_t me, int64_t exp10) {
  if (a.nsig == 0) return 0.0;
  const int64_t e = a.dropped - a.frac + exp10;   // value = w x 10^e when exact
  if (a.exact && a.w < ((uint64_t)1 << 53)) {
#if defined(__CUDA_ARCH__)
    if (e >= 0 && e <= 22) return __dmul_rn((double)a.w, pow10_exact((int)e));
    if (e < 0 && e >= -22) return __ddiv_rn((double)a.w, pow10_exact((int)-e));
#else
    if (e >= 0 && e <= 22) return (double)a.w * pow10_exact((int)e);
    if (e < 0 && e >= -22) return (double)a.w / pow10_exact((int)-e);
#endif
  }
  return big_to_double(t, ms, me, exp10);
}

// value of mantissa text [ms, me) x 10^exp10 (unsigned), correctly rounded.
O3V_HD double dec_to_double(const uint8_t* t, int64_t ms, int64_t me, int64_t exp10) {
  DecAcc a;
  acc_init(a);
  bool seen_point = false;
  for (int64_t p = ms; p < me;) {
    const uint8_t c = t[p];
    if (c == '.') { seen_point = true; ++p; continue; }
    if (c == '_') { ++p; continue; }
    int len = 1;
    const int v = digit_at(t, p, me, &len);
    p += len;
    if (v >= 0) acc_digit(a, v, seen_point);
  }
  return acc_value(a, t, ms, me, exp10);
}

// ------------------------------------------------------------------------------------------
// Number grammars.
// ------------------------------------------------------------------------------------------
// `[\d.]+` at p (reward_func.py:405, :447): returns the end of the run (p if empty); *ok says
// whether float() accepts it (>= 1 digit, <= 1 dot).
O3V_HD int64_t scan_digits_dots(const uint8_t* t, int64_t p, int64_t end, bool* ok, DecAcc* acc) {
  int digits = 0, dots = 0;
  acc_init(*acc);
  while (p < end) {
    if (t[p] == '.') { ++dots; ++p; continue; }
    int len;
    const int v = digit_at(t, p, end, &len);
    if (v < 0) break;
    acc_digit(*acc, v, dots > 0);
    ++digits;
    p += len;
  }
  *ok = digits >= 1 && dots <= 1;
  return p;
}
// `\d+\.?\d*` at p (reward_func.py:119): end of the match, or -1.
O3V_HD int64_t scan_simple_decimal(const uint8_t* t, int64_t p, int64_t end) {
  int len, n = 0;
  while (p < end && digit_at(t, p, end, &len) >= 0) { p += len; ++n; }
  if (n == 0) return -1;
  if (p < end && t[p] == '.') ++p;
  while (p < end && digit_at(t, p, end, &len) >= 0) p += len;
  return p;
}

O3V_HD uint8_t lower_ascii(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }
O3V_HD bool word_at_ci(const uint8_t* t, int64_t p, int64_t end, const Lit& l) {
  if (p + l.n > end) return false;
  for (int i = 0; i < l.n; ++i)
    if (lower_ascii(t[p + i]) != lit_byte(l, i)) return false;
  return true;
}

// digits with optional single '_' between digits; returns end or -1 if no digit at p.
O3V_HD int64_t scan_digitpart(const uint8_t* t, int64_t p, int64_t end, bool* any) {
  int len;
  *any = false;
  if (digit_at(t, p, end, &len) < 0) return p;
  p += len;
  *any = true;
  for (;;) {
    if (p < end && t[p] == '_') {
      if (digit_at(t, p + 1, end, &len) < 0) return -1;    // '_' must sit between two digits
      p += 1 + len;
      continue;
    }
    if (digit_at(t, p, end, &len) < 0) break;
    p += len;
  }
  return p;
}

// explicit exponent digits -> saturated int64
O3V_HD int64_t exp_value(const uint8_t* t, int64_t s, int64_t e) {
  int64_t v = 0;
  for (int64_t p = s; p < e;) {
    if (t[p] == '_') { ++p; continue; }
    int len = 1;
    const int d = digit_at(t, p, e, &len);
    p += len;
    if (d < 0) continue;
    if (v < ((int64_t)1 << 40)) v = v * 10 + d;
  }
  return v;
}

// Python float(str) on the (already stripped) range [s, e): reward_func.py:318.
O3V_HD bool python_float(const uint8_t* t, int64_t s, int64_t e, double* out) {
  constexpr Lit kInf = make_lit("inf"), kInfinity = make_lit("infinity"), kNan = make_lit("nan");
  int64_t p = s;
  bool neg = false;
  if (p < e && (t[p] == '+' || t[p] == '-')) { neg = t[p] == '-'; ++p; }
  if (p >= e) return false;
  const uint8_t c = lower_ascii(t[p]);
  if (c == 'i' || c == 'n') {
    double v;
    if (word_at_ci(t, p, e, kInfinity) && p + 8 == e) v = bits_to_double(0x7FF0000000000000ull);
    else if (word_at_ci(t, p, e, kInf) && p + 3 == e) v = bits_to_double(0x7FF0000000000000ull);
    else if (word_at_ci(t, p, e, kNan) && p + 3 == e) v = bits_to_double(0x7FF8000000000000ull);
    else return false;
    *out = neg ? -v : v;
    return true;
  }
  const int64_t ms = p;
  bool any_int = false, any_frac = false;
  p = scan_digitpart(t, p, e, &any_int);
  if (p < 0) return false;
  if (p < e && t[p] == '.') {
    ++p;
    p = scan_digitpart(t, p, e, &any_frac);
    if (p < 0) return false;
  }
  if (!any_int && !any_frac) return false;
  const int64_t me = p;
  int64_t exp10 = 0;
  if (p < e && (t[p] == 'e' || t[p] == 'E')) {
    ++p;
    bool eneg = false;
    if (p < e && (t[p] == '+' || t[p] == '-')) { eneg = t[p] == '-'; ++p; }
    bool any_exp = false;
    const int64_t es = p;
    p = scan_digitpart(t, p, e, &any_exp);
    if (p < 0 || !any_exp) return false;
    exp10 = exp_value(t, es, p);
    if (eneg) exp10 = -exp10;
  }
  if (p != e) return false;
  const double v = dec_to_double(t, ms, me, exp10);
  *out = neg ? -v : v;
  return true;
}

// ------------------------------------------------------------------------------------------
// JSON recogniser for a box payload (json.loads of CPython's C scanner, reward_func.py:216, :223,
// :322, :511).  The payload starts with '['.  Result:
//   kJsonInvalid : json.loads raises (JSONDecodeError)
//   otherwise    : *n_elems = len(list); *numeric = np.array(list, dtype=float) succeeds
//                  (reward_func.py:364-365): numbers, true / false (1.0 / 0.0), null and NaN (nan),
//                  +-Infinity, and strings that float() accepts; elements 0..3 are in out[].
// Other strings, objects and ragged nested lists raise inside np.array and are caught (:367):
// `numeric = false`, the box scores 0.  Not reproduced: strings with backslash
// escapes that unescape to a float (treated as non-numeric) and homogeneous nested lists
// ([[1],[2],[3],[4]]: a 2-D array in the reference, which then fails on an ambiguous truth value).
// ------------------------------------------------------------------------------------------
constexpr int kJsonInvalid = 0, kJsonList = 1;
constexpr int kJsonMaxDepth = 64;

O3V_HD bool json_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
O3V_HD bool is_hex(uint8_t c) {
  return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F');
}

// string body after the opening quote; returns position after the closing quote or -1
O3V_HD int64_t json_string(const uint8_t* t, int64_t p, int64_t e) {
  while (p < e) {
    const uint8_t c = t[p];
    if (c == '"') return p + 1;
    if (c < 0x20) return -1;                               // strict: no control characters
    if (c == '\\') {
      if (p + 1 >= e) return -1;
      const uint8_t x = t[p + 1];
      if (x == 'u') {
        if (p + 6 > e) return -1;
        for (int i = 2; i < 6; ++i)
          if (!is_hex(t[p + i])) return -1;
        p += 6;
      } else if (x == '"' || x == '\\' || x == '/' || x == 'b' || x == 'f' || x == 'n' || x == 'r' || x == 't') {
        p += 2;
      } else {
        return -1;
      }
      continue;
    }
    ++p;
  }
  return -1;
}

// JSON number at p: -?(0|[1-9]\d*)(\.\d+)?([eE][-+]?\d+)?  -> end, value; -1 if none
O3V_HD int64_t json_number(const uint8_t* t, int64_t p, int64_t e, double* out) {
  bool neg = false;
  if (p < e && t[p] == '-') { neg = true; ++p; }
  const int64_t ms = p;
  if (p >= e || t[p] < '0' || t[p] > '9') return -1;
  DecAcc a;
  acc_init(a);
  if (t[p] == '0') ++p;
  else
    for (; p < e && t[p] >= '0' && t[p] <= '9'; ++p) acc_digit(a, t[p] - '0', false);
  bool is_float = false;
  if (p + 1 < e && t[p] == '.' && t[p + 1] >= '0' && t[p + 1] <= '9') {
    is_float = true;
    for (++p; p < e && t[p] >= '0' && t[p] <= '9'; ++p) acc_digit(a, t[p] - '0', true);
  }
  const int64_t me = p;
  int64_t exp10 = 0;
  if (p < e && (t[p] == 'e' || t[p] == 'E')) {
    int64_t q = p + 1;
    bool eneg = false;
    if (q < e && (t[q] == '+' || t[q] == '-')) { eneg = t[q] == '-'; ++q; }
    if (q < e && t[q] >= '0' && t[q] <= '9') {
      const int64_t es = q;
      while (q < e && t[q] >= '0' && t[q] <= '9') ++q;
      exp10 = exp_value(t, es, q);
      if (eneg) exp10 = -exp10;
      is_float = true;
      p = q;
    }
  }
  const double v = acc_value(a, t, ms, me, exp10);
  // an integer literal becomes a Python int: "-0" is int 0 -> +0.0 in np.array(dtype=float)
  *out = (neg && (is_float || v != 0.0)) ? -v : v;
  return p;
}

O3V_HD int json_box(const uint8_t* t, int64_t s, int64_t e, int* n_elems, bool* numeric, double out[4],
                    unsigned lanes = 0xffffffffu) {
  constexpr Lit kNull = make_lit("null"), kTrue = make_lit("true"), kFalse = make_lit("false"),
                kNaN = make_lit("NaN"), kInfinity = make_lit("Infinity"), kNegInfinity = make_lit("-Infinity");
  int64_t p = s;
  uint64_t is_array = 0;     // bit d: container at depth d+1 is an array
  int depth = 0;
  int n_top = 0;
  bool all_numeric = true;
  // states
  enum { kValueOrClose, kValue, kAfterValue, kKeyOrClose, kKey, kColon, kDone, kBad };
  int state = kValue;
  // one token per iteration; the lanes of a warp (each on its own payload) start every token together
  O3V_LOCKSTEP_BEGIN(lanes, state < kDone)
    while (p < e && json_ws(t[p])) ++p;
    if (p >= e) { state = kBad; continue; }
    const uint8_t c = t[p];
    if (state == kValue || state == kValueOrClose) {
      if (state == kValueOrClose && c == ']') {            // empty array
        --depth; ++p;
        state = depth == 0 ? kDone : kAfterValue;
        if (depth == 1) { ++n_top; all_numeric = false; }  // a nested [] element
        continue;
      }
      const bool top_elem = depth == 1;
      if (c == '[' || c == '{') {
        if (depth >= kJsonMaxDepth) { state = kBad; continue; }
        if (c == '[') is_array |= (uint64_t)1 << depth; else is_array &= ~((uint64_t)1 << depth);
        ++depth; ++p;
        state = c == '[' ? kValueOrClose : kKeyOrClose;
        continue;                                          // the element is counted when it closes
      }
      double v = 0.0;
      bool num = false;
      if (c == '"') {
        const int64_t q = json_string(t, p + 1, e);
        if (q < 0) { state = kBad; continue; }
        if (top_elem) {                                      // np.array(dtype=float) calls float(str)
          int64_t fs = p + 1, fe = q - 1;
          bool escaped = false;
          for (int64_t i = fs; i < fe; ++i) escaped |= t[i] == '\\';
          strip_space(t, &fs, &fe, false);
          num = !escaped && python_float(t, fs, fe, &v);
        }
        p = q;
      } else if (c == 'n' && lit_at(t, p, e, kNull)) { p += 4; v = bits_to_double(0x7FF8000000000000ull); num = true;
      } else if (c == 't' && lit_at(t, p, e, kTrue)) { p += 4; v = 1.0; num = true;
      } else if (c == 'f' && lit_at(t, p, e, kFalse)) { p += 5; v = 0.0; num = true;
      } else if (c == 'N' && lit_at(t, p, e, kNaN)) { p += 3; v = bits_to_double(0x7FF8000000000000ull); num = true;
      } else if (c == 'I' && lit_at(t, p, e, kInfinity)) { p += 8; v = bits_to_double(0x7FF0000000000000ull); num = true;
      } else if (c == '-' && lit_at(t, p, e, kNegInfinity)) { p += 9; v = bits_to_double(0xFFF0000000000000ull); num = true;
      } else {
        const int64_t q = json_number(t, p, e, &v);
        if (q < 0) { state = kBad; continue; }
        p = q;
        num = true;
      }
      if (top_elem) {
        if (num && n_top < 4) out[n_top] = v;
        if (!num) all_numeric = false;
        ++n_top;
      }
      state = kAfterValue;
      continue;
    }
    if (state == kAfterValue) {
      const bool arr = (is_array >> (depth - 1)) & 1;
      if (c == ',') { ++p; state = arr ? kValue : kKey; continue; }
      if ((arr && c == ']') || (!arr && c == '}')) {
        --depth; ++p;
        if (depth == 0) { state = kDone; continue; }
        if (depth == 1) { ++n_top; all_numeric = false; }  // a nested container element closed
        state = kAfterValue;
        continue;
      }
      { state = kBad; continue; }
    }
    if (state == kKeyOrClose || state == kKey) {
      if (state == kKeyOrClose && c == '}') {
        --depth; ++p;
        if (depth == 0) { state = kDone; continue; }       // not reachable: the payload starts with '['
        if (depth == 1) { ++n_top; all_numeric = false; }
        state = kAfterValue;
        continue;
      }
      if (c != '"') { state = kBad; continue; }
      p = json_string(t, p + 1, e);
      if (p < 0) { state = kBad; continue; }
      state = kColon;
      continue;
    }
    if (state == kColon) {
      if (c != ':') { state = kBad; continue; }
      ++p;
      state = kValue;
      continue;
    }
  O3V_LOCKSTEP_END
  if (state == kBad) return kJsonInvalid;
  while (p < e && json_ws(t[p])) ++p;
  if (p != e) return kJsonInvalid;                          // "Extra data"
  *n_elems = n_top;
  *numeric = all_numeric;
  return kJsonList;
}

// ------------------------------------------------------------------------------------------
// One rollout, in three phases.
// ------------------------------------------------------------------------------------------
struct Caps { int P, C, Bc, Tb; };
struct RolloutOut {            // pointers to THIS rollout's rows (o3v_rewards_soa layout)
  int32_t* flags; double* ans_seg; double* ans_box; int32_t* n_times; double* think_times;
  int32_t* n_claims; double* claim_t; int32_t* claim_nbox; uint32_t* claim_valid; double* claim_box;
  int32_t* n_tboxes; uint32_t* tbox_valid; double* think_box;
};
// Per-rollout scratch between the phases (caller-provided workspace, 40 bytes per rollout).
struct Scratch {
  int64_t think_end, answer_end;
  int32_t time_cands, claim_cands, tbox_cands;
  uint32_t unused0_, unused1_;        // (were 32-bit kept / numeric masks of the think boxes: now slot markers)
  int32_t pad_;
};

#if defined(__CUDA_ARCH__)
#define O3V_LANE0 ((threadIdx.x & 31) == 0)
#else
#define O3V_LANE0 true
#endif

constexpr int kFlagThink = 1, kFlagAnswer = 2, kFlagAnsSeg = 4, kFlagAnsBox = 8;
constexpr int kTaskVisual = 0, kTaskTemporal = 1, kTaskTemporalMcq = 2;

// A candidate is a byte range of the text, stored in the double slot its value will occupy:
// start in the high 40 bits, length in the low 24 (a completion is shorter than 16 MiB).
constexpr int64_t kMaxRange = (1 << 24) - 1;
O3V_HD double pack_range(int64_t s, int64_t e) {
  int64_t n = e - s;
  if (n > kMaxRange) n = kMaxRange;       // cannot be a number; rejected in phase B
  return bits_to_double(((uint64_t)s << 24) | (uint64_t)n);
}
O3V_HD void unpack_range(double d, int64_t* s, int64_t* e) {
#if defined(__CUDA_ARCH__)
  const uint64_t u = (uint64_t)__double_as_longlong(d);
#else
  union { double d; uint64_t u; } c;
  c.d = d;
  const uint64_t u = c.u;
#endif
  *s = (int64_t)(u >> 24);
  *e = *s + (int64_t)(u & 0xFFFFFF);
}
// phase B -> C status of a think-time candidate (real values are >= 0 or +inf)
constexpr double kTimeNoMatch = -1.0, kTimeBadFloat = -2.0;

// `<box>(\[.*?\])</box>` without DOTALL at or after p inside [.., lim): on success the payload is
// [*bs, *be) (brackets included) and the return value is the end of the match; -1 if no match.
// '.' does not cross a newline: a candidate whose first "]</box>" lies behind a newline fails, and so
// does every candidate before that newline.
O3V_HD int64_t next_box(Finder& f, int64_t p, int64_t lim, int64_t* bs, int64_t* be) {
  for (;;) {
    p = f.find<kLBoxOpen>(p, lim);
    if (p < 0) return -1;
    const int64_t c = f.find<kLBoxClose>(p + 6, lim);
    if (c < 0) return -1;
    const int64_t nl = f.find<kLNewline>(p + 6, c);
    if (nl < 0) {
      *bs = p + 5;
      *be = c + 1;
      return c + 7;
    }
    p = nl + 1;
  }
}

// ---- phase A: spans and match chains -> candidates (device: warp-uniform, lane 0 stores)
O3V_HD void scan_rollout(const uint8_t* t, int64_t total, int64_t beg, int64_t end, int task, const Caps& cap,
                         const RolloutOut& o, Scratch* sc, uint32_t* warp_smem) {
  constexpr Lit kTEnd = make_lit("</t>s"), kTo = make_lit("</t>s to <t>");
  Finder f;
  f.init(t, total, beg, warp_smem);
  int flags = 0;
  // think / answer spans: leftmost open tag, first close tag after it (lazy `.*?`, DOTALL)
  int64_t ts = f.find<kLThinkO>(beg, end), te = -1;
  if (ts >= 0) { ts += 7; te = f.find<kLThinkC>(ts, end); }
  const bool has_think = te >= 0;
  int64_t as = f.find<kLAnsO>(beg, end), ae = -1;
  if (as >= 0) { as += 8; ae = f.find<kLAnsC>(as, end); }
  const bool has_answer = ae >= 0;
  if (has_think) flags |= kFlagThink;
  if (has_answer) flags |= kFlagAnswer;

  // ---- answer "<t>a</t>s to <t>b</t>s" (:119, temporal tasks): the first <t> at which the whole
  // pattern matches; the two number ranges are converted in phase B
  if ((task == kTaskTemporal || task == kTaskTemporalMcq) && has_answer) {
    int64_t p = as;
    while ((p = f.find<kLT>(p, ae)) >= 0) {
      const int64_t a0 = p + 3, a1 = scan_simple_decimal(t, a0, ae);
      if (a1 >= 0 && lit_at(t, a1, ae, kTo)) {
        const int64_t b0 = a1 + 12, b1 = scan_simple_decimal(t, b0, ae);
        if (b1 >= 0 && lit_at(t, b1, ae, kTEnd)) {
          flags |= kFlagAnsSeg;
          if (O3V_LANE0) { o.ans_seg[0] = pack_range(a0, a1); o.ans_seg[1] = pack_range(b0, b1); }
          break;
        }
      }
      ++p;
    }
  }

  // ---- every "<t>" inside <think> is a candidate for "<t>x</t>s" (:405-412, :447-449); a match
  // contains no second "<t>", so the candidates of findall's matches are exactly these, in order
  int n_time = 0, n_tbox = 0, n_claim = 0;
  if (has_think) {
#if defined(__CUDA_ARCH__)
    // all occurrences at once: per 512-byte block every lane owns the matches among its 16 positions; a warp
    // prefix sum of the match counts gives each lane the candidate index of its first match
    const int lane = threadIdx.x & 31;
    const int64_t last = te - 3;
    for (int64_t blk = ts & ~(int64_t)511; blk <= last; blk += 512) {
      uint32_t m = f.block_mask<kLT>(blk, ts, last);
      const int cnt = __popc(m);
      int pre = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += v;
      }
      int k = n_time + pre - cnt;
      for (; m; m &= m - 1, ++k)
        if (k < cap.P) {
          const int64_t x0 = blk + lane * 16 + (__ffs(m) - 1) + 3;
          o.think_times[k] = pack_range(x0, x0);
        }
      n_time += __shfl_sync(0xffffffffu, pre, 31);
    }
#else
    int64_t p = ts;
    while ((p = f.find<kLT>(p, te)) >= 0) {
      if (n_time < cap.P) o.think_times[n_time] = pack_range(p + 3, p + 3);
      ++n_time;
      p += 3;
    }
#endif
  }

  if (task == kTaskVisual) {
    // ---- first <box> of the answer (:211-223); flag provisional until phase B has parsed it
    if (has_answer) {
      int64_t bs, be;
      if (next_box(f, as, ae, &bs, &be) >= 0) {
        flags |= kFlagAnsBox;
        if (O3V_LANE0) o.ans_box[0] = pack_range(bs, be);
      }
    }
    // ---- every <box> inside <think> (:505-511)
    if (has_think) {
      int64_t p = ts, bs, be;
      while ((p = next_box(f, p, te, &bs, &be)) >= 0) {
        if (n_tbox < cap.Tb && O3V_LANE0) o.think_box[4 * n_tbox] = pack_range(bs, be);
        ++n_tbox;
      }
    }
  } else if (has_think) {
    // ---- claims (:308-335): <obj>(.*?)</obj>((?:<box>\[.*?\]</box>)+)at<t>(.*?)</t>s, DOTALL.
    // Leftmost match at an <obj>: shortest group 1 = first "</obj><box>[" after it; group 2 ends at
    // the first "]</box>at<t>" after that; group 3 at the first "</t>s" (DESIGN.md section 8).
#if defined(__CUDA_ARCH__)
    // Lane-parallel form of the same chain when the span's masks fit the shared-memory cache: all <obj>
    // occurrences of a block are listed at once, lane i looks up (q, e, z) of occurrence i on its own, and
    // only the choice "first <obj> at or after the previous match's end" runs serially (on shuffles).
    bool fast = false;
    const int lane = threadIdx.x & 31;
    const int64_t b0 = (ts - f.first) >> 9, b1 = (te - f.first) >> 9;
    if (b1 < Finder::kSmemBlocks) {
      fast = true;
      for (int64_t b = b0; b <= b1; ++b) (void)f.block_mask<kLObj>(f.first + (b << 9), ts, te);
      __syncwarp();
      uint32_t* list = f.sm + Finder::kMaskWords;
      int64_t cur = ts;
      bool stop = false;
      for (int64_t b = b0; b <= b1 && !stop; ++b) {
        const int64_t blk = f.first + (b << 9);
        uint32_t m = f.block_mask<kLObj>(blk, ts, te - 5);
        const int cnt = __popc(m);
        int pre = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, pre, d);
          if (lane >= d) pre += v;
        }
        const int total_objs = __shfl_sync(0xffffffffu, pre, 31);
        if (total_objs == 0) continue;
        if (total_objs > 32) { fast = false; break; }          // pathological block: take the sequential chain
        for (int k = pre - cnt; m; m &= m - 1, ++k) list[k] = (uint32_t)(blk + lane * 16 + (__ffs(m) - 1) - f.first);
        __syncwarp();
        int64_t p = -1, q = -1, e = -1, z = -1;
        if (lane < total_objs) {
          p = f.first + list[lane];
          q = f.next_lane<kLObjBox>(p + 5, te);
          if (q >= 0) e = f.next_lane<kLBoxAt>(q + 12, te);
          if (e >= 0) z = f.next_lane<kLTEnd>(e + 12, te);
        }
        __syncwarp();
        for (int i = 0; i < total_objs; ++i) {
          const int64_t pi = __shfl_sync(0xffffffffu, p, i);
          if (pi < cur) continue;                              // inside the previous match
          const int64_t zi = __shfl_sync(0xffffffffu, z, i);
          if (zi < 0) { stop = true; break; }                  // a failed search ends finditer
          if (lane == i && n_claim < cap.C) {
            o.claim_t[n_claim] = pack_range(e + 12, z);
            o.claim_box[(int64_t)n_claim * cap.Bc * 4] = pack_range(q + 6, e + 7);
          }
          ++n_claim;
          cur = zi + 5;
        }
      }
      if (!fast) n_claim = 0;
    }
    if (!fast)
#endif
    {
    int64_t p = ts;
    for (;;) {
      p = f.find<kLObj>(p, te);
      if (p < 0) break;
      const int64_t q = f.find<kLObjBox>(p + 5, te);
      if (q < 0) break;
      const int64_t e = f.find<kLBoxAt>(q + 12, te);
      if (e < 0) break;
      const int64_t z = f.find<kLTEnd>(e + 12, te);
      if (z < 0) break;
      p = z + 5;
      if (n_claim < cap.C && O3V_LANE0) {
        o.claim_t[n_claim] = pack_range(e + 12, z);                            // group 3
        o.claim_box[(int64_t)n_claim * cap.Bc * 4] = pack_range(q + 6, e + 7);  // group 2
      }
      ++n_claim;
    }
    }
  }

  if (O3V_LANE0) {
    *o.flags = flags;
    sc->think_end = te;
    sc->answer_end = ae;
    sc->time_cands = n_time;
    sc->claim_cands = n_claim;
    sc->tbox_cands = n_tbox;
    sc->unused0_ = 0;
    sc->unused1_ = 0;
  }
}

// ---- phase B: one candidate (device: one thread each).  Items of a rollout are numbered
// [0, P) think times, [P, P+C) claims, [P+C, P+C+Tb) think boxes, then the answer segment and the
// answer box.
O3V_HD int items_per_rollout(const Caps& cap) { return cap.P + cap.C + cap.Tb + 2; }

O3V_HD void or_bits(uint32_t* word, uint32_t bits) {
#if defined(__CUDA_ARCH__)
  atomicOr(word, bits);
#else
  *word |= bits;
#endif
}

// does this rollout hold candidate `item`?  (the lanes for which it does run convert_item together)
O3V_HD bool item_active(int item, const Caps& cap, const RolloutOut& o, const Scratch* sc) {
  if (item < cap.P) return item < sc->time_cands;
  item -= cap.P;
  if (item < cap.C) return item < sc->claim_cands;
  item -= cap.C;
  if (item < cap.Tb) return item < sc->tbox_cands;
  item -= cap.Tb;
  return (*o.flags & (item == 0 ? kFlagAnsSeg : kFlagAnsBox)) != 0;
}

O3V_HD void convert_item(const uint8_t* t, int item, const Caps& cap, const RolloutOut& o, Scratch* sc,
                         unsigned lanes) {
  constexpr Lit kTEnd = make_lit("</t>s");
  if (item < cap.P) {                                                   // ---- "<t>x</t>s" of <think>
    if (item >= sc->time_cands) return;
    int64_t x0, unused;
    unpack_range(o.think_times[item], &x0, &unused);
    const int64_t te = sc->think_end;
    bool ok;
    DecAcc acc;
    const int64_t x1 = scan_digits_dots(t, x0, te, &ok, &acc);
    double v = kTimeNoMatch;
    if (x1 > x0 && lit_at(t, x1, te, kTEnd)) v = ok ? acc_value(acc, t, x0, x1, 0) : kTimeBadFloat;
    o.think_times[item] = v;
    return;
  }
  item -= cap.P;
  if (item < cap.C) {                                                   // ---- one claim
    if (item >= sc->claim_cands) return;
    int64_t fs, fe, g0, g1;
    unpack_range(o.claim_t[item], &fs, &fe);
    double* cbox = o.claim_box + (int64_t)item * cap.Bc * 4;
    unpack_range(cbox[0], &g0, &g1);
    const bool too_long = fe - fs >= kMaxRange || g1 - g0 >= kMaxRange;
    strip_space(t, &fs, &fe, true);
    double tv;
    int nb = 0;
    uint32_t valid = 0;
    bool keep = !too_long && python_float(t, fs, fe, &tv);               // ValueError -> claim dropped (:332)
    // boxes: re.findall(r'\[.*?\]', group 2) without DOTALL
    int64_t i = g0;
    O3V_SYNC_LANES(lanes);
    O3V_LOCKSTEP_BEGIN(lanes, keep && i < g1)                           // one box per iteration
      while (i < g1 && t[i] != '[') ++i;
      int64_t j = i + 1;
      int found = 0;                                                     // 1: payload [i, j], 2: a newline came first
      if (i < g1) {
        while (j < g1 && t[j] != ']' && t[j] != '\n') ++j;
        found = j >= g1 ? 0 : (t[j] == ']' ? 1 : 2);
      }
      const unsigned parsers = O3V_LOCKSTEP_WHO(found == 1);
      if (found == 1) {
        int n; bool numeric; double v[4];
        if (json_box(t, i, j + 1, &n, &numeric, v, parsers) != kJsonList) { keep = false; continue; }   // JSONDecodeError
        // validity of box b: bit b of `valid` for b < 32; beyond that (degenerate repetition loops) an invalid
        // box is marked IN its slot with kInvalidBoxBits, a NaN payload no parsed number can have
        if (n == 4 && numeric) {
          if (nb < 32) valid |= 1u << nb;
          if (nb < cap.Bc) {
            double* dst = cbox + 4 * nb;
            dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3];
          }
        } else if (nb >= 32 && nb < cap.Bc) {
          cbox[4 * nb] = bits_to_double(kInvalidBoxBits);
        }
        ++nb;
        i = j + 1;
      } else {
        i = found == 2 ? j + 1 : g1;
      }
    O3V_LOCKSTEP_END
    o.claim_nbox[item] = keep ? nb : -1;
    if (keep) {
      o.claim_t[item] = tv;
      o.claim_valid[item] = valid;
    }
    return;
  }
  item -= cap.C;
  if (item < cap.Tb) {                                                  // ---- one <box> of <think> (visual QA)
    if (item >= sc->tbox_cands) return;
    int64_t bs, be;
    double* dst = o.think_box + 4 * item;
    unpack_range(dst[0], &bs, &be);
    int n; bool numeric; double v[4];
    if (json_box(t, bs, be - bs >= kMaxRange ? bs : be, &n, &numeric, v, lanes) != kJsonList) {         // not JSON: skipped
      dst[0] = bits_to_double(kDroppedBoxBits);
      return;
    }
    // kept (JSON list): a 4-number box keeps its values, anything else is marked invalid in its slot; a payload
    // that is not JSON was marked dropped below.  (Slot markers instead of per-rollout bit masks: no 32-box limit.)
    if (n == 4 && numeric) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3]; }
    else dst[0] = bits_to_double(kInvalidBoxBits);
    return;
  }
  item -= cap.Tb;
  if (item == 0) {                                                      // ---- the answer's two numbers
    if (!(*o.flags & kFlagAnsSeg)) return;
    for (int k = 0; k < 2; ++k) {
      int64_t s, e;
      unpack_range(o.ans_seg[k], &s, &e);
      o.ans_seg[k] = dec_to_double(t, s, e, 0);
    }
    return;
  }
  if (!(*o.flags & kFlagAnsBox)) return;                                // ---- the answer's first <box>
  int64_t bs, be;
  unpack_range(o.ans_box[0], &bs, &be);
  int n; bool numeric; double v[4];
  if (json_box(t, bs, be - bs >= kMaxRange ? bs : be, &n, &numeric, v, lanes) == kJsonList && n == 4 && numeric) {
    o.ans_box[0] = v[0]; o.ans_box[1] = v[1]; o.ans_box[2] = v[2]; o.ans_box[3] = v[3];
  } else {
    *o.flags &= ~kFlagAnsBox;
  }
}

// ---- phase C: drop rejected candidates, compact, publish the counts.  over[4]: a count that did not
// fit its capacity (candidates are an upper bound of the matches), else 0.
O3V_HD void finish_rollout(const Caps& cap, const RolloutOut& o, const Scratch* sc, int over[4]) {
  over[0] = sc->time_cands > cap.P ? sc->time_cands : 0;
  over[1] = sc->claim_cands > cap.C ? sc->claim_cands : 0;
  over[2] = 0;
  over[3] = sc->tbox_cands > cap.Tb ? sc->tbox_cands : 0;
  // think times: one bad float() empties the list (:413-415, :450-451)
  int n = sc->time_cands < cap.P ? sc->time_cands : cap.P, k = 0;
  bool all_ok = true;
  for (int i = 0; i < n; ++i) {
    const double v = o.think_times[i];
    if (v == kTimeBadFloat) all_ok = false;
    else if (v != kTimeNoMatch) { if (k != i) o.think_times[k] = v; ++k; }
  }
  *o.n_times = all_ok ? k : 0;
  // claims
  n = sc->claim_cands < cap.C ? sc->claim_cands : cap.C;
  k = 0;
  for (int c = 0; c < n; ++c) {
    const int nb = o.claim_nbox[c];
    if (nb < 0) continue;
    if (nb > cap.Bc && nb > over[2]) over[2] = nb;
    if (k != c) {
      o.claim_t[k] = o.claim_t[c];
      o.claim_nbox[k] = nb;
      o.claim_valid[k] = o.claim_valid[c];
      const int m = (nb < cap.Bc ? nb : cap.Bc) * 4;
      const double* src = o.claim_box + (int64_t)c * cap.Bc * 4;
      double* dst = o.claim_box + (int64_t)k * cap.Bc * 4;
      for (int i = 0; i < m; ++i) dst[i] = src[i];
    }
    ++k;
  }
  *o.n_claims = k;
  // think boxes
  n = sc->tbox_cands < cap.Tb ? sc->tbox_cands : cap.Tb;
  k = 0;
  uint32_t valid = 0;
  for (int j = 0; j < n; ++j) {
    const uint64_t b0 = double_to_bits(o.think_box[4 * j]);
    if (b0 == kDroppedBoxBits) continue;
    if (b0 != kInvalidBoxBits) {
      if (k < 32) valid |= 1u << k;
      if (k != j)
        for (int i = 0; i < 4; ++i) o.think_box[4 * k + i] = o.think_box[4 * j + i];
    } else if (k >= 32) {
      o.think_box[4 * k] = bits_to_double(kInvalidBoxBits);
    }
    ++k;
  }
  *o.n_tboxes = k;
  *o.tbox_valid = valid;
}

}  // namespace scan
}  // namespace o3v
