// exact_sum_kernels.cuh -- Field.sum (Field.pde:311-318) as three kernels around exact_sum.cuh:
//   k_xsum_totals  double-precision sum of every chunk (kXsChunk additions) of every environment
//   k_xsum_tables  one thread per segment of 32 additions: predicted accumulator before the segment (chunk
//                  totals + block scan), then the segment summary (exact_sum.cuh) -> 64 B slot
//   k_xsum_chain   one warp per environment walks the slots with the true accumulator; a segment whose summary
//                  does not provably apply is redone as 32 genuine float additions
// The result is bit-identical to the serial loop for any input (tests/test_exact_sum.py, tests/test_gpu_parity.py).
// Included by solver_kernels.cu inside namespace rlfc::{anonymous}.
#pragma once
// (exact_sum.cuh is included at file scope by solver_kernels.cu)

constexpr int kXsThreads = 256;
constexpr int kXsChunk = kXsThreads * xsum::kSeg;       // additions per CTA
constexpr int kXsPad = xsum::kSeg + 1;                  // shared-memory stride of a segment (bank-conflict free)
constexpr int kXsRecWords = 64;                         // batch record: 8 header words + kXsRecEntries entries of 8 words
constexpr int kXsRecEntries = 7;

// serial index K (i-major over the interior) -> offset in the pitched array.  advance() is branch-free when a
// row is at least half a CTA wide (WIDE: at most two row ends per stride of kXsThreads elements), so that the
// loads of an unrolled loop are not separated by control flow.
template <bool WIDE>
struct XsCursor {
  int i, j, len, P;
  __device__ __forceinline__ XsCursor(long long K, int len_, int P_) : len(len_), P(P_) {
    const unsigned k = (unsigned)K;                               // rlfc_env_create rejects grids beyond 2^31 cells
    i = 1 + (int)(k / (unsigned)len_); j = 1 + (int)(k % (unsigned)len_);
  }
  __device__ __forceinline__ size_t off() const { return (size_t)i * P + j; }
  __device__ __forceinline__ void advance() {
    j += kXsThreads;
    if (WIDE) {
      const bool a = j > len; j -= a ? len : 0; i += a;
      const bool b = j > len; j -= b ? len : 0; i += b;
    } else {
      while (j > len) { j -= len; i++; }
    }
  }
};

template <bool WIDE>
__device__ __forceinline__ double xs_chunk_total(const SolverParams& q, const float* p, long long base, long long N, int t) {
  double s = 0.0;
  XsCursor<WIDE> cur(base + t, q.m - 2, q.P);
  const unsigned left = (unsigned)(N - base);                     // elements of this chunk and beyond (N > base)
#pragma unroll 1
  for (int h = 0; h < xsum::kSeg; h += 8) {                       // 8 independent loads in flight per thread
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      v[u] = ((unsigned)(t + (h + u) * kXsThreads) < left) ? p[cur.off()] : 0.f;
      cur.advance();
    }
#pragma unroll
    for (int u = 0; u < 8; u++) s += (double)v[u];
  }
  return s;
}

__global__ void __launch_bounds__(kXsThreads)
k_xsum_totals(const __grid_constant__ SolverParams q) {
  __shared__ double wsum[kXsThreads / 32];
  const int c = blockIdx.x, e = blockIdx.y, t = threadIdx.x;
  const int len = q.m - 2;
  const long long N = (long long)(q.n - 2) * len, base = (long long)c * kXsChunk;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  double s = (2 * len >= kXsThreads) ? xs_chunk_total<true>(q, p, base, N, t) : xs_chunk_total<false>(q, p, base, N, t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((t & 31) == 0) wsum[t >> 5] = s;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int w = 0; w < kXsThreads / 32; w++) tot += wsum[w];
    q.xs_ctot[(size_t)e * q.xs_nchunks + c] = tot;
  }
}

template <bool WIDE>
__device__ __forceinline__ void xs_fill(const SolverParams& q, const float* p, long long base, long long N, int t, float* buf) {
  XsCursor<WIDE> cur(base + t, q.m - 2, q.P);
#pragma unroll 8
  for (int u = 0; u < xsum::kSeg; u++) {
    const int kl = t + u * kXsThreads;                            // local index: segment kl / 32, element kl % 32
    buf[(kl >> 5) * kXsPad + (kl & 31)] = (base + kl < N) ? p[cur.off()] : 0.f;
    cur.advance();
  }
}

__global__ void __launch_bounds__(kXsThreads)
k_xsum_tables(const __grid_constant__ SolverParams q) {
  __shared__ float buf[kXsThreads * kXsPad];
  __shared__ double wsum[kXsThreads / 32];
  __shared__ double base_pred;
  const int c = blockIdx.x, e = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int len = q.m - 2;
  const long long N = (long long)(q.n - 2) * len, base = (long long)c * kXsChunk;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  if (2 * len >= kXsThreads) xs_fill<true>(q, p, base, N, t, buf);
  else xs_fill<false>(q, p, base, N, t, buf);
  if (t == 0) {
    double b = 0.0;
    const double* ct = q.xs_ctot + (size_t)e * q.xs_nchunks;
    for (int k = 0; k < c; k++) b += ct[k];
    base_pred = b;
  }
  __syncthreads();
  const long long g = (long long)c * kXsThreads + t;              // this thread's segment
  const long long left = N - g * xsum::kSeg;
  const int cnt = left >= xsum::kSeg ? xsum::kSeg : (left > 0 ? (int)left : 0);
  const float* seg = buf + t * kXsPad;
  float fsum = 0.f;                                               // a prediction only: float precision is plenty
#pragma unroll 8
  for (int k = 0; k < xsum::kSeg; k++) fsum += seg[k];            // (tail elements are stored as +0)
  const double ssum = (double)fsum;
  // exclusive scan of the segment sums over the CTA
  double incl = ssum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  double pred = base_pred;
  for (int w = 0; w < warp; w++) pred += wsum[w];
  pred += incl - ssum;
  uint32_t slot[xsum::kSlotWords];
  if (cnt > 0) {
    xsum::build_segment([&](int k) { return seg[k]; }, cnt, pred, slot);
  } else {                                                        // past the end: an empty table
#pragma unroll
    for (int w = 0; w < xsum::kSlotWords; w++) slot[w] = 0u;
    slot[0] = xsum::kOne; slot[1] = xsum::kAnyKey;
  }
  // Batch record: the 32 summaries of this warp (= one batch of the serial pass) condensed into a few entries
  // "table, then one float addition".  A split or serial summary (a "head") closes the running group of tables:
  //   C_k = table of the group that is open after slot k  (head: what it restarts with -- the second table of
  //         a split summary, nothing for a serial one; otherwise C_{k-1} followed by the slot's table),
  //   E_k = C_{k-1} followed by the head's first table = the entry that ends at head k, with the head's float
  //         addition (split) or the 32 genuine additions of the segment (serial) right behind it;
  // the last entry is C_31.  C is a Kogge-Stone scan of table compositions, segmented at the heads.
  {
    const uint32_t type = slot[0];
    const bool head = type != xsum::kOne;
    uint32_t X[7], acc[7];
    const uint32_t ident[7] = {xsum::kAnyKey, 0u, 0u, (uint32_t)INT32_MIN, (uint32_t)INT32_MAX, (uint32_t)INT32_MIN, (uint32_t)INT32_MAX};
    xsum::normalise_table(slot + 1);
    if (type == xsum::kSplit) xsum::normalise_table(slot + 9);
#pragma unroll
    for (int w = 0; w < 7; w++) {
      X[w] = (type == xsum::kSerial) ? ident[w] : slot[1 + w];
      acc[w] = (type == xsum::kOne) ? slot[1 + w] : (type == xsum::kSplit ? slot[9 + w] : ident[w]);
    }
    const uint32_t headmask = __ballot_sync(0xffffffffu, head);
    const uint32_t upto = headmask & ((2u << lane) - 1u);         // heads at or below this lane
    const int dist = lane - (upto ? 31 - __clz(upto) : 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t prev[7];
#pragma unroll
      for (int w = 0; w < 7; w++) prev[w] = __shfl_up_sync(0xffffffffu, acc[w], o);
      if (dist >= o) xsum::compose_tables(prev, acc);
    }
    uint32_t cprev[7];
#pragma unroll
    for (int w = 0; w < 7; w++) {
      cprev[w] = __shfl_up_sync(0xffffffffu, acc[w], 1);
      if (lane == 0) cprev[w] = ident[w];
    }
    xsum::compose_tables(cprev, X);                               // X = E_k on head lanes
    const int idx = __popc(headmask & ((1u << lane) - 1u));       // entry index of a head lane
    const int count = __popc(headmask) + 1;
    const uint32_t serbits = __reduce_or_sync(0xffffffffu, type == xsum::kSerial ? (1u << idx) : 0u);
    const long long batch = (long long)c * (kXsThreads / 32) + warp;
    if (batch < q.xs_nbatches) {
      uint32_t* rec = q.xs_recs + ((size_t)e * q.xs_nbatches + batch) * kXsRecWords;
      if (lane == 0)
        *reinterpret_cast<uint4*>(rec) = make_uint4(count <= kXsRecEntries ? (uint32_t)count : 0xffffffffu, serbits, headmask, 0u);
      if (count <= kXsRecEntries) {
        if (head) {
          uint4* ent = reinterpret_cast<uint4*>(rec + 8 + 8 * idx);
          ent[0] = make_uint4(X[0], X[1], X[2], X[3]);
          ent[1] = make_uint4(X[4], X[5], X[6], type == xsum::kSplit ? slot[8] : 0x80000000u);   // -0.f: s + -0 == s
        }
        if (lane == 31) {
          uint4* ent = reinterpret_cast<uint4*>(rec + 8 + 8 * (count - 1));
          ent[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
          ent[1] = make_uint4(acc[4], acc[5], acc[6], 0x80000000u);
        }
      }
    }
  }
  if (g < q.xs_nseg) {
    uint4* out = reinterpret_cast<uint4*>(q.xs_slots + ((size_t)e * q.xs_nseg + g) * xsum::kSlotWords);
#pragma unroll
    for (int w = 0; w < 4; w++) out[w] = make_uint4(slot[4 * w], slot[4 * w + 1], slot[4 * w + 2], slot[4 * w + 3]);
  }
}

// Serial pass.  One warp per environment walks the batch records: per entry one table application (checked) and
// one float addition, plus the 32 genuine additions of a serial segment where an entry says so.  If any entry
// of a batch does not provably apply (or the batch has too many heads for a record), the batch is redone from
// its start summary by summary (walk): chain of  bits += (bits & 1) ? D1 : D0  over the plain summaries, every
// lane checks its own summary's condition, the first one that fails is redone as 32 float additions.
constexpr int kXsRing = 4;        // batch records in the shared-memory ring (cp.async, 3 batches ahead)
constexpr int kXsRawPf = 4;       // serial segments per batch whose elements are prefetched

__device__ __forceinline__ void xs_cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(32)
k_xsum_chain(const __grid_constant__ SolverParams q) {
  __shared__ __align__(16) uint32_t ring[kXsRing][kXsRecWords];
  const int e = blockIdx.x, lane = threadIdx.x;
  const unsigned len = (unsigned)(q.m - 2), P = (unsigned)q.P;
  const unsigned N = (unsigned)(q.n - 2) * len;                   // rlfc_env_create rejects grids beyond 2^31 cells
  const int nseg = q.xs_nseg, nb = q.xs_nbatches;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  const uint4* slots = reinterpret_cast<const uint4*>(q.xs_slots + (size_t)e * nseg * xsum::kSlotWords);
  const uint32_t* recs = q.xs_recs + (size_t)e * nb * kXsRecWords;
  auto element = [&](unsigned g) {                                 // this lane's element of segment g (0 past the end)
    const unsigned K = g * xsum::kSeg + lane;
    return K < N ? p[(size_t)(1u + K / len) * P + 1u + K % len] : 0.f;
  };
  auto fetch = [&](int b) {                                       // one commit group per batch, also when empty
    if (b < nb && lane < kXsRecWords / 4) xs_cp16(ring[b % kXsRing] + 4 * lane, recs + (size_t)b * kXsRecWords + 4 * lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // xs_stats[e][0..3]: batches crossed by their record / walked summary by summary, record entries applied,
  // segments redone as float additions
  int st_rec = 0, st_walk = 0, st_ent = 0, st_redo = 0;
  uint32_t bits = 0u;                                             // s = +0.f
  // 32 genuine additions of segment g; lane u holds element u in v
  auto redo = [&](unsigned g, float v) {
    st_redo++;
    const unsigned K0 = g * xsum::kSeg;
    const int cnt = K0 >= N ? 0 : (N - K0 >= (unsigned)xsum::kSeg ? xsum::kSeg : (int)(N - K0));
    float el[xsum::kSeg];
#pragma unroll
    for (int u = 0; u < xsum::kSeg; u++) el[u] = __shfl_sync(0xffffffffu, v, u);
    float s = xsum::u2f(bits);
#pragma unroll
    for (int u = 0; u < xsum::kSeg; u++) s = (u < cnt) ? s + el[u] : s;
    bits = xsum::f2u(s);
  };
  // batch b summary by summary, from the accumulator in `bits`
  auto walk = [&](int b) {
    st_walk++;
    uint32_t w[16];
    {
      const int g = b * 32 + lane;
      uint4 v4[4];
      if (g < nseg) {
#pragma unroll
        for (int k = 0; k < 4; k++) v4[k] = slots[(size_t)g * 4 + k];
      } else {                                                    // past the end: an empty table
        v4[0] = make_uint4(xsum::kOne, xsum::kAnyKey, 0u, 0u);
        v4[1] = v4[2] = v4[3] = make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) { w[4 * k] = v4[k].x; w[4 * k + 1] = v4[k].y; w[4 * k + 2] = v4[k].z; w[4 * k + 3] = v4[k].w; }
    }
    const uint32_t special = __ballot_sync(0xffffffffu, w[0] != xsum::kOne);
    int cur = 0;
    while (cur < 32) {
      const uint32_t rest = special >> cur;
      const int f = rest ? cur + __ffs(rest) - 1 : 32;            // next summary that is not a plain table
      uint32_t acc = bits, mine = bits;
      for (int k0 = cur; k0 < f; k0++) {
        const uint32_t d0 = __shfl_sync(0xffffffffu, w[2], k0), d1 = __shfl_sync(0xffffffffu, w[3], k0);
        mine = (lane == k0) ? acc : mine;
        acc += (acc & 1u) ? d1 : d0;
      }
      bool ok = true;
      if (lane >= cur && lane < f)
        xsum::apply_table(mine, w[1], (int32_t)w[2], (int32_t)w[3], (int32_t)w[4], (int32_t)w[5], (int32_t)w[6], (int32_t)w[7], ok);
      const uint32_t bad = __ballot_sync(0xffffffffu, !ok);
      int ev;
      if (bad) { ev = __ffs(bad) - 1; bits = __shfl_sync(0xffffffffu, mine, ev); }
      else { ev = f; bits = acc; }
      if (ev < 32) {
        uint32_t sw[16];
#pragma unroll
        for (int k = 0; k < 16; k++) sw[k] = __shfl_sync(0xffffffffu, w[k], ev);
        if (bad || !xsum::apply_segment(bits, sw)) redo((unsigned)(b * 32 + ev), element((unsigned)(b * 32 + ev)));
      }
      cur = ev + 1;
    }
  };
  fetch(0); fetch(1); fetch(2);
  float rawn[kXsRawPf];                                           // elements of the next batch's first serial segments
#pragma unroll
  for (int i = 0; i < kXsRawPf; i++) rawn[i] = 0.f;
  for (int b = 0; b < nb; b++) {
    fetch(b + 3);
    asm volatile("cp.async.wait_group 2;" ::: "memory");          // records <= b + 1 have landed
    __syncwarp();
    const uint32_t* rec = ring[b % kXsRing];
    const uint4 hdr = *reinterpret_cast<const uint4*>(rec);       // count, serial entries, head summaries
    float raw[kXsRawPf];
#pragma unroll
    for (int i = 0; i < kXsRawPf; i++) raw[i] = rawn[i];
    if (b + 1 < nb) {                                             // elements of the next batch's serial segments
      const uint4 nh = *reinterpret_cast<const uint4*>(ring[(b + 1) % kXsRing]);
      uint32_t m = nh.x == 0xffffffffu ? 0u : nh.y;
#pragma unroll
      for (int i = 0; i < kXsRawPf; i++) {
        rawn[i] = 0.f;
        if (m) {
          const int ent = __ffs(m) - 1;
          m &= m - 1u;
          rawn[i] = element((unsigned)((b + 1) * 32) + __fns(nh.z, 0u, ent + 1));
        }
      }
    }
    if (hdr.x == 0xffffffffu) { walk(b); continue; }
    const uint32_t start = bits;
    const int st_redo0 = st_redo;
    bool ok = true;
    const int count = (int)hdr.x;
    for (int i = 0; i < count; i++) {
      const uint4 t0 = *reinterpret_cast<const uint4*>(rec + 8 + 8 * i);
      const uint4 t1 = *reinterpret_cast<const uint4*>(rec + 12 + 8 * i);
      bits = xsum::apply_table(bits, t0.x, (int32_t)t0.y, (int32_t)t0.z, (int32_t)t0.w, (int32_t)t1.x, (int32_t)t1.y,
                               (int32_t)t1.z, ok);
      bits = xsum::f2u(xsum::u2f(bits) + xsum::u2f(t1.w));
      if ((hdr.y >> i) & 1u) {                                    // the entry ends at a serial segment
        const unsigned g = (unsigned)(b * 32) + __fns(hdr.z, 0u, i + 1);
        const int si = __popc(hdr.y & ((1u << i) - 1u));
        float v;
        if (b > 0 && si < kXsRawPf) {
          v = raw[0];
#pragma unroll
          for (int k = 1; k < kXsRawPf; k++) v = (si == k) ? raw[k] : v;
        } else {
          v = element(g);                                         // (batch 0 has no prefetch)
        }
        redo(g, v);
      }
    }
    if (ok) { st_rec++; st_ent += count; }
    else { bits = start; st_redo = st_redo0; walk(b); }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (lane == 0) {
    q.sc.psum[e] = xsum::u2f(bits);
    int* st = q.xs_stats + 8 * e;
    st[0] = st_rec; st[1] = st_walk; st[2] = st_ent; st[3] = st_redo;
  }
}
