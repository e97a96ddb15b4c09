// exact_sum.cuh -- Field.sum (Field.pde:311-318) evaluated in parallel, bit-identical to the serial loop.
//
// The reference adds the interior of p into ONE float accumulator in i-major order:  s = fl(s + a_k).  The
// result depends on every intermediate rounding, so it cannot be re-associated -- but it can be re-PHRASED.
// While the accumulator stays inside one binade [2^e, 2^(e+1)) with a fixed sign, it is an integer S (its
// 24-bit significand) times u = 2^(e-23), and  fl(s + a) = u * RN(S + a/u)  with RN = round to nearest
// integer, ties to even.  Writing  a/u = +-(Q + rem/2^sh)  (Q, rem from the significand of a, sh = e - e_a),
//     S' = S +- (Q + r),   r = 0 if rem/2^sh < 1/2,  1 if > 1/2,  parity(S) ^ (Q & 1) on an exact tie,
// i.e. a run of additions is an INTEGER sum whose only dependence on the running value is one parity bit
// (and after the first tie the parity is even whatever it was before).  A run is therefore summarised by a
// table { D[h], lo[h], hi[h] : h = parity of S at the start }:  S_end = S + D[h], valid iff lo[h] <= S <= hi[h]
// (every intermediate significand stays strictly inside (2^23, 2^24), which is exactly the condition under
// which the integer model equals IEEE-754 addition) and the accumulator really has the sign and exponent the
// table was built for.  Tables of different segments are built independently -- in parallel -- from a
// PREDICTED accumulator (double-precision prefix sums of the data, accurate to ~1e-5 relative, far finer than
// a binade); the addition at which the predicted accumulator changes binade or sign, and its two neighbours, are
// kept as genuine float additions between two tables.  Tables compose (associatively), so a run of segments
// condenses into one table.  A last, short serial pass walks the condensed summaries with the true accumulator,
// checks every table's validity condition, and recomputes whatever fails (wrong prediction, several binade
// changes, accumulator near zero, Inf/NaN) with the plain serial loop.  Nothing is approximate: a table is only
// applied when it provably reproduces the serial additions bit for bit.
//
// This header holds the host/device core (segment summary, composition, application); tests/test_exact_sum.py
// drives it on the CPU against the serial loop (tests/xsum_host.cpp), exact_sum_kernels.cuh wraps it in two kernels.
#pragma once
#include <climits>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define XS_HD __host__ __device__ __forceinline__
#else
#define XS_HD inline
#endif

namespace rlfc {
namespace xsum {

constexpr int kSeg = 32;                 // additions per segment summary
constexpr int kSlotWords = 20;           // summary: [type | table (7) | 3 floats | table (7) | 2 unused]
constexpr int kSlotRaw = 8, kSlotB = 11; // word offsets of the float additions and of the second table
constexpr uint32_t kNegZero = 0x80000000u;   // s + -0.f == s for every s: the "no addition" float
constexpr uint32_t kOne = 0;             // one table covers the whole segment
constexpr uint32_t kSplit = 1;           // table, three genuine float additions (around the binade/sign change), table
constexpr uint32_t kSerial = 2;          // no summary: the serial pass adds the segment's elements one by one
constexpr uint32_t kAnyKey = 0xffffffffu;   // table of an empty run: applies to any accumulator, changes nothing

XS_HD uint32_t f2u(float v) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(v);
#else
  uint32_t u; std::memcpy(&u, &v, 4); return u;
#endif
}
XS_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float v; std::memcpy(&v, &u, 4); return v;
#endif
}

// key of an accumulator value = sign and biased exponent (bits 31..23).  Tables exist for normal values with
// biased exponent in [30, 254]: below that the significand arithmetic of tiny/denormal addends would need
// more cases than it is worth (such accumulators go through the serial path).
XS_HD uint32_t key_of(float v) { return f2u(v) >> 23; }
XS_HD bool key_ok(uint32_t key) { const uint32_t e = key & 255u; return e >= 30u && e <= 254u; }

// A table is seven words: [key, D0, D1, lo0, hi0, lo1, hi1].

// Running summary of a run of additions for an accumulator with sign/exponent `key`.
// The integer increment of one addition is read off the FPU instead of being assembled from shifts:  with
// R = 1.5 * 2^e (significand 0xC00000, made odd when the modelled significand is odd) and b = the addend with
// the accumulator's sign factored out,  fl(R + b) - R  is exactly  u * (+-(Q + r))  as long as R + b stays
// inside the binade -- guaranteed for |b| < 2^(e-2) -- because the rounding of R + b sees the same fraction
// and the same significand parity as the real accumulator.  In the same binade the difference of the two bit
// patterns IS that integer.  Larger addends (within a factor 8 of the accumulator) are not modelled.
struct Run {
  uint32_t key, sgnmask, Rb, lim;
  int32_t P0, P1, mn0, mn1, mx0, mx1;
  int good, empty;                       // (ints, not bools: the compiler packs bools into bytes and pays PRMTs for it)
  XS_HD void start(uint32_t key_) {
    key = key_;
    const uint32_t es = key & 255u;
    sgnmask = (key >> 8) << 31;
    Rb = (es << 23) | 0x400000u;
    lim = (es - 2u) << 23;                         // |addend| must be below 2^(e-2)
    P0 = P1 = 0; mn0 = mn1 = INT32_MAX; mx0 = mx1 = INT32_MIN;
    good = 1; empty = 1;
  }
  XS_HD void add(float a) {
    const uint32_t ab = f2u(a);
    good &= (int)((ab & 0x7fffffffu) < lim);       // also rejects Inf / NaN
    const float b = u2f(ab ^ sgnmask);
    const uint32_t r0 = Rb | ((uint32_t)P0 & 1u), r1 = Rb | (((uint32_t)P1 & 1u) ^ 1u);
    P0 += (int32_t)(f2u(u2f(r0) + b) - r0);
    P1 += (int32_t)(f2u(u2f(r1) + b) - r1);
    mn0 = P0 < mn0 ? P0 : mn0; mx0 = P0 > mx0 ? P0 : mx0;
    mn1 = P1 < mn1 ? P1 : mn1; mx1 = P1 > mx1 ? P1 : mx1;
    empty = 0;
  }
  // words [key, D0, D1, lo0, hi0, lo1, hi1]; every significand after an addition must stay in
  // [2^23 + 1, 2^24 - 1] (header comment)
  XS_HD void store(uint32_t* w) const {
    if (empty) {
      w[0] = kAnyKey; w[1] = w[2] = 0; w[3] = w[5] = (uint32_t)INT32_MIN; w[4] = w[6] = (uint32_t)INT32_MAX;
      return;
    }
    w[0] = key; w[1] = (uint32_t)P0; w[2] = (uint32_t)P1;
    w[3] = (uint32_t)((int32_t)0x800001 - mn0); w[4] = (uint32_t)((int32_t)0xffffff - mx0);
    w[5] = (uint32_t)((int32_t)0x800001 - mn1); w[6] = (uint32_t)((int32_t)0xffffff - mx1);
  }
};

// Apply a table to the accumulator bits; clears `ok` when the table does not provably apply.
XS_HD uint32_t apply_table(uint32_t bits, uint32_t key, int32_t D0, int32_t D1, int32_t lo0, int32_t hi0, int32_t lo1,
                           int32_t hi1, bool& ok) {
  if (key == kAnyKey) return bits;
  const bool odd = (bits & 1u) != 0;
  const int32_t S = (int32_t)((bits & 0x7fffffu) | 0x800000u);
  const int32_t lo = odd ? lo1 : lo0, hi = odd ? hi1 : hi0;
  ok = ok && (bits >> 23) == key && S >= lo && S <= hi;
  return bits + (uint32_t)(odd ? D1 : D0);
}

// Summary of one segment of cnt <= kSeg additions, given the predicted accumulator before its first addition.
// slot[kSlotWords]: see the constants above.  One pass: the prediction is carried forward as a float chain
// (what the real accumulator would do from the predicted start).  The addition across which the prediction
// changes sign or binade (or before which it has no usable key) cannot be part of a table; it is kept as a
// genuine float addition TOGETHER WITH ITS TWO NEIGHBOURS, because the real accumulator differs from the
// prediction by a few ulps and may change binade one addition earlier or later.  With one such place the
// segment is  table | 3 float additions | table;  with more it is left to the serial pass.
template <typename Get>
XS_HD void build_segment(Get get, int cnt, double pred_start, uint32_t* slot) {
  float pred = (float)pred_start;
  uint32_t key_prev = key_of(pred);
  Run run;
  run.start(key_prev);
  int state = 0;                         // 0: first table, 1: the addition after the change is due, 2: second table
  int good = 1, pending = 0, serial = 0;
  // additions enter a table one step late (`pend`), so that the one before a change can still be kept out of it
  float pend = 0.f;
  int force = key_ok(key_prev) ? 0 : 1;  // non-zero: the next addition takes the rare path whatever its key
  for (int w = 0; w < kSlotWords; w++) slot[w] = 0;
  slot[kSlotRaw] = slot[kSlotRaw + 1] = slot[kSlotRaw + 2] = kNegZero;
  for (int k = 0; k < cnt; k++) {
    const float a = get(k);
    pred = pred + a;
    const uint32_t kn = key_of(pred);
    if (kn == key_prev && !force) {      // common path (threads of a warp stay converged here)
      if (pending) run.add(pend);
      pend = a; pending = 1;
      continue;
    }
    if (state == 0) {                    // the change: close the first table WITHOUT the previous addition
      good = run.good | run.empty;
      run.store(slot + 1);
      slot[kSlotRaw] = pending ? f2u(pend) : kNegZero;
      slot[kSlotRaw + 1] = f2u(a);
      pending = 0;
      state = 1; force = 1;
    } else if (state == 1) {             // a genuine addition whatever it does; the second table starts behind it
      slot[kSlotRaw + 2] = f2u(a);
      run.start(kn);
      if (!key_ok(kn)) run.good = 0;     // only acceptable if nothing follows (checked through `empty`)
      state = 2; force = 0;
    } else {
      serial = 1;                        // a second change: left to the serial pass
    }
    key_prev = kn;
  }
  if (pending) run.add(pend);
  if (state == 0) {
    if (!run.good) { slot[0] = kSerial; return; }
    slot[0] = kOne;
    run.store(slot + 1);
    return;
  }
  if (state == 1) { run.start(key_prev); run.good = 1; }        // nothing behind the change: empty second table
  good = good & (serial ^ 1) & (run.good | run.empty);
  if (!good) { slot[0] = kSerial; return; }
  slot[0] = kSplit;
  run.store(slot + kSlotB);
}

// ---- composition of tables (used to summarise a whole stretch of plain segments in one table) ----
// A table half (one start parity) that can never apply is normalised to D = 0, lo = kNever, hi = -kNever, so
// that composing it further cannot overflow or turn it valid again.
constexpr int32_t kNever = 0x40000000;
XS_HD void normalise_half(int32_t& D, int32_t& lo, int32_t& hi) {
  // a table that applies moves the significand by less than 2^24; lo > hi is the empty interval
  if (lo > hi || D >= (int32_t)0x1000000 || D <= -(int32_t)0x1000000) { D = 0; lo = kNever; hi = -kNever; }
}
// w = [key, D0, D1, lo0, hi0, lo1, hi1]
XS_HD void normalise_table(uint32_t* w) {
  if (w[0] == kAnyKey) return;
  int32_t D0 = (int32_t)w[1], D1 = (int32_t)w[2], lo0 = (int32_t)w[3], hi0 = (int32_t)w[4], lo1 = (int32_t)w[5],
          hi1 = (int32_t)w[6];
  normalise_half(D0, lo0, hi0);
  normalise_half(D1, lo1, hi1);
  w[1] = (uint32_t)D0; w[2] = (uint32_t)D1; w[3] = (uint32_t)lo0; w[4] = (uint32_t)hi0; w[5] = (uint32_t)lo1; w[6] = (uint32_t)hi1;
}
// one start parity of (f then g): the half of g that follows is chosen by the parity f leaves behind
XS_HD void compose_half(bool same, int32_t fD, int32_t flo, int32_t fhi, int32_t gD0, int32_t gD1, int32_t glo0, int32_t ghi0,
                        int32_t glo1, int32_t ghi1, int32_t& D, int32_t& lo, int32_t& hi) {
  const bool odd = (fD & 1) != 0;
  const int32_t gD = odd ? gD1 : gD0, glo = odd ? glo1 : glo0, ghi = odd ? ghi1 : ghi0;
  const bool dead = !same || flo == kNever || glo == kNever;
  const int32_t l2 = glo - fD, u2 = ghi - fD;
  D = fD + gD;
  lo = flo > l2 ? flo : l2;
  hi = fhi < u2 ? fhi : u2;
  if (dead) { lo = kNever; hi = -kNever; }
  normalise_half(D, lo, hi);
}
// g <- (f then g): the additions of f followed by those of g.  Both normalised; the result is normalised.
XS_HD void compose_tables(const uint32_t* f, uint32_t* g) {
  if (f[0] == kAnyKey) return;
  if (g[0] == kAnyKey) { for (int k = 0; k < 7; k++) g[k] = f[k]; return; }
  const bool same = f[0] == g[0];                    // a table only follows another inside the same binade
  const int32_t gD0 = (int32_t)g[1], gD1 = (int32_t)g[2], glo0 = (int32_t)g[3], ghi0 = (int32_t)g[4], glo1 = (int32_t)g[5],
                ghi1 = (int32_t)g[6];
  int32_t D0, lo0, hi0, D1, lo1, hi1;
  compose_half(same, (int32_t)f[1], (int32_t)f[3], (int32_t)f[4], gD0, gD1, glo0, ghi0, glo1, ghi1, D0, lo0, hi0);
  compose_half(same, (int32_t)f[2], (int32_t)f[5], (int32_t)f[6], gD1, gD0, glo1, ghi1, glo0, ghi0, D1, lo1, hi1);
  g[0] = f[0];
  g[1] = (uint32_t)D0; g[2] = (uint32_t)D1; g[3] = (uint32_t)lo0; g[4] = (uint32_t)hi0; g[5] = (uint32_t)lo1; g[6] = (uint32_t)hi1;
}

// Advance the accumulator over one summarised segment.  Returns true when the summary applied; false = the
// caller must redo the segment serially from `bits` (which is left untouched in that case).
XS_HD bool apply_segment(uint32_t& bits, const uint32_t* slot) {
  if (slot[0] == kSerial) return false;
  bool ok = true;
  uint32_t b = apply_table(bits, slot[1], (int32_t)slot[2], (int32_t)slot[3], (int32_t)slot[4], (int32_t)slot[5],
                           (int32_t)slot[6], (int32_t)slot[7], ok);
  if (slot[0] == kSplit) {
    b = f2u(((u2f(b) + u2f(slot[kSlotRaw])) + u2f(slot[kSlotRaw + 1])) + u2f(slot[kSlotRaw + 2]));
    b = apply_table(b, slot[kSlotB], (int32_t)slot[kSlotB + 1], (int32_t)slot[kSlotB + 2], (int32_t)slot[kSlotB + 3],
                    (int32_t)slot[kSlotB + 4], (int32_t)slot[kSlotB + 5], (int32_t)slot[kSlotB + 6], ok);
  }
  if (ok) bits = b;
  return ok;
}

}  // namespace xsum
}  // namespace rlfc
