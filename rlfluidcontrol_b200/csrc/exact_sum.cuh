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
// a binade); the one addition at which the predicted accumulator changes binade or sign is kept as a genuine
// float addition between two tables.  A last, short serial pass walks the segment summaries with the true
// accumulator, checks every table's validity condition, and recomputes any segment that fails (wrong
// prediction, several binade changes, accumulator near zero, Inf/NaN) with the plain serial loop.  Nothing is
// approximate: a table is only applied when it provably reproduces the serial additions bit for bit.
//
// This header holds the host/device core (segment summary + application); tests/test_exact_sum.py drives it
// on the CPU against the serial loop, solver_kernels.cu wraps it in three kernels.
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
constexpr int kSlotWords = 16;           // summary size: [type | table (7) | raw float | table (7)]
constexpr uint32_t kOne = 0;             // one table covers the whole segment
constexpr uint32_t kSplit = 1;           // table, one genuine float addition (the binade/sign change), table
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

struct Table {
  uint32_t key;
  int32_t D0, D1, lo0, hi0, lo1, hi1;
};

// Summarise the additions a[k0 .. k1) for an accumulator with sign/exponent `key`.  `get(k)` returns element k.
// Returns false when some addend cannot be modelled (Inf/NaN, or an addend at least as large as the
// accumulator's binade: the sum necessarily leaves the binade).
template <typename Get>
XS_HD bool build_table(Get get, int k0, int k1, uint32_t key, Table& T) {
  if (k1 <= k0) {
    T.key = kAnyKey; T.D0 = T.D1 = 0; T.lo0 = T.lo1 = INT32_MIN; T.hi0 = T.hi1 = INT32_MAX;
    return true;
  }
  const uint32_t sgn = key >> 8, es = key & 255u;
  int32_t P0 = 0, P1 = 0, mn0 = INT32_MAX, mn1 = INT32_MAX, mx0 = INT32_MIN, mx1 = INT32_MIN;
  uint32_t c0 = 0, c1 = 1;                      // parity of the significand under the two start hypotheses
  bool good = true;
  for (int k = k0; k < k1; k++) {
    const uint32_t ab = f2u(get(k));
    const bool neg = (((ab >> 31) ^ sgn) & 1u) != 0;      // sign of the addend relative to the accumulator
    uint32_t ea = (ab >> 23) & 255u, M = ab & 0x7fffffu;
    if (ea == 255u) good = false;
    if (ea) M |= 0x800000u; else ea = 1;                    // denormal / zero: exponent -126, no hidden bit
    const int sh = (int)es - (int)ea;
    if (sh < 1) { good = false; continue; }
    uint32_t Q, gt, tie;                                     // a/u = Q + rem/2^sh; gt: rem > half, tie: rem == half
    if (sh > 24) { Q = 0; gt = 0; tie = 0; }
    else {
      Q = M >> sh;
      const uint32_t rem = M & ((1u << sh) - 1u), half = 1u << (sh - 1);
      gt = rem > half; tie = rem == half;
    }
    const uint32_t r0 = gt | (tie & (c0 ^ Q)), r1 = gt | (tie & (c1 ^ Q));
    const int32_t i0 = (int32_t)(Q + (r0 & 1u)), i1 = (int32_t)(Q + (r1 & 1u));
    P0 += neg ? -i0 : i0;
    P1 += neg ? -i1 : i1;
    c0 ^= (uint32_t)i0; c1 ^= (uint32_t)i1;                 // only bit 0 is used
    mn0 = P0 < mn0 ? P0 : mn0; mx0 = P0 > mx0 ? P0 : mx0;
    mn1 = P1 < mn1 ? P1 : mn1; mx1 = P1 > mx1 ? P1 : mx1;
  }
  T.key = key; T.D0 = P0; T.D1 = P1;
  // every significand after an addition must stay in [2^23 + 1, 2^24 - 1] (header comment)
  T.lo0 = (int32_t)0x800001 - mn0; T.hi0 = (int32_t)0xffffff - mx0;
  T.lo1 = (int32_t)0x800001 - mn1; T.hi1 = (int32_t)0xffffff - mx1;
  return good;
}

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
// slot[kSlotWords]: see the constants above.
template <typename Get>
XS_HD void build_segment(Get get, int cnt, double pred, uint32_t* slot) {
  // States 0..cnt = predicted accumulator before element k.  Element k cannot be part of a table when the
  // prediction changes sign or binade across it (or has no usable key before it): with at most one such
  // element the segment is  table [0,kx) | float addition of a[kx] | table (kx,cnt).
  const uint32_t key_first = key_of((float)pred);
  uint32_t key_prev = key_first, key_after = key_first;
  int ncross = 0, kx = 0;
  for (int k = 0; k < cnt; k++) {
    pred += (double)get(k);
    const uint32_t kn = key_of((float)pred);
    if (kn != key_prev || !key_ok(key_prev)) { ncross++; kx = k; key_after = kn; }
    key_prev = kn;
  }
  for (int w = 0; w < kSlotWords; w++) slot[w] = 0;
  Table A, B;
  bool good = ncross <= 1;
  if (good && ncross == 0) {
    good = build_table(get, 0, cnt, key_first, A);
    if (good) {
      slot[0] = kOne;
      slot[1] = A.key; slot[2] = (uint32_t)A.D0; slot[3] = (uint32_t)A.D1; slot[4] = (uint32_t)A.lo0;
      slot[5] = (uint32_t)A.hi0; slot[6] = (uint32_t)A.lo1; slot[7] = (uint32_t)A.hi1;
      return;
    }
  } else if (good) {
    good = build_table(get, 0, kx, key_first, A) && build_table(get, kx + 1, cnt, key_after, B);
    if (good) {
      slot[0] = kSplit;
      slot[1] = A.key; slot[2] = (uint32_t)A.D0; slot[3] = (uint32_t)A.D1; slot[4] = (uint32_t)A.lo0;
      slot[5] = (uint32_t)A.hi0; slot[6] = (uint32_t)A.lo1; slot[7] = (uint32_t)A.hi1;
      slot[8] = f2u(get(kx));
      slot[9] = B.key; slot[10] = (uint32_t)B.D0; slot[11] = (uint32_t)B.D1; slot[12] = (uint32_t)B.lo0;
      slot[13] = (uint32_t)B.hi0; slot[14] = (uint32_t)B.lo1; slot[15] = (uint32_t)B.hi1;
      return;
    }
  }
  slot[0] = kSerial;
}

// Advance the accumulator over one summarised segment.  Returns true when the summary applied; false = the
// caller must redo the segment serially from `bits` (which is left untouched in that case).
XS_HD bool apply_segment(uint32_t& bits, const uint32_t* slot) {
  if (slot[0] == kSerial) return false;
  bool ok = true;
  uint32_t b = apply_table(bits, slot[1], (int32_t)slot[2], (int32_t)slot[3], (int32_t)slot[4], (int32_t)slot[5],
                           (int32_t)slot[6], (int32_t)slot[7], ok);
  if (slot[0] == kSplit) {
    b = f2u(u2f(b) + u2f(slot[8]));
    b = apply_table(b, slot[9], (int32_t)slot[10], (int32_t)slot[11], (int32_t)slot[12], (int32_t)slot[13],
                    (int32_t)slot[14], (int32_t)slot[15], ok);
  }
  if (ok) bits = b;
  return ok;
}

}  // namespace xsum
}  // namespace rlfc
