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
  // Stretch tables: inside the warp (= one batch of 32 slots of the serial pass), every plain slot also gets the
  // composition of all plain tables from the start of its stretch up to itself (words 8..14, free in a plain
  // slot), so the serial pass crosses a whole stretch with one table.  Kogge-Stone scan, segmented at the
  // non-plain slots.
  {
    const bool plain = slot[0] == xsum::kOne;
    if (plain) xsum::normalise_table(slot + 1);
    const uint32_t pm = __ballot_sync(0xffffffffu, plain);
    // plain slots directly below this lane: distance to the start of the stretch
    const uint32_t below = ~pm & ((1u << lane) - 1u);
    const int start = below ? 32 - __clz(below) : 0;
    const int dist = lane - start;
    uint32_t acc[7];
#pragma unroll
    for (int w = 0; w < 7; w++) acc[w] = slot[1 + w];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t prev[7];
#pragma unroll
      for (int w = 0; w < 7; w++) prev[w] = __shfl_up_sync(0xffffffffu, acc[w], o);
      if (plain && dist >= o) xsum::compose_tables(prev, acc);
    }
    if (plain) {
#pragma unroll
      for (int w = 0; w < 7; w++) slot[8 + w] = acc[w];
    }
  }
  if (g < q.xs_nseg) {
    uint4* out = reinterpret_cast<uint4*>(q.xs_slots + ((size_t)e * q.xs_nseg + g) * xsum::kSlotWords);
#pragma unroll
    for (int w = 0; w < 4; w++) out[w] = make_uint4(slot[4 * w], slot[4 * w + 1], slot[4 * w + 2], slot[4 * w + 3]);
  }
}

// Serial pass.  One warp per environment; lane L holds summary 32 b + L of batch b (registers + a shared-memory
// copy for broadcast reads).  A stretch of plain slots is crossed with the stretch table of its last slot; a
// slot that is not plain (split / serial) is an event handled on its own.  If a stretch table does not apply,
// the stretch is walked slot by slot (chain of  bits += (bits & 1) ? D1 : D0,  then all lanes check their own
// slot's condition at once) and the first slot whose own table fails is redone as 32 float additions.
constexpr int kXsRing = 4;        // batches of summaries in the shared-memory ring (cp.async, 3 batches ahead)
constexpr int kXsRawPf = 4;       // serial slots per batch whose elements are prefetched

__device__ __forceinline__ void xs_cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(32)
k_xsum_chain(const __grid_constant__ SolverParams q) {
  __shared__ __align__(16) uint32_t ring[kXsRing][32 * xsum::kSlotWords];
  const int e = blockIdx.x, lane = threadIdx.x;
  const unsigned len = (unsigned)(q.m - 2), P = (unsigned)q.P;
  const unsigned N = (unsigned)(q.n - 2) * len;                   // rlfc_env_create rejects grids beyond 2^31 cells
  const int nseg = q.xs_nseg;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  const uint4* slots = reinterpret_cast<const uint4*>(q.xs_slots + (size_t)e * nseg * xsum::kSlotWords);
  auto element = [&](unsigned g) {                                 // this lane's element of segment g (0 past the end)
    const unsigned K = g * xsum::kSeg + lane;
    return K < N ? p[(size_t)(1u + K / len) * P + 1u + K % len] : 0.f;
  };
  const int nb = (nseg + 31) / 32;
  // batch b -> ring slot b % kXsRing; every lane copies its own summary (one commit group per batch, also when empty)
  auto fetch = [&](int b) {
    if (b < nb) {
      uint32_t* dst = ring[b % kXsRing] + lane * xsum::kSlotWords;
      const int g = b * 32 + lane;
      if (g < nseg) {                                             // (lanes past the last summary are never looked at)
#pragma unroll
        for (int w = 0; w < 4; w++) xs_cp16(dst + 4 * w, slots + (size_t)g * 4 + w);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // genuine serial additions of one segment (warp-uniform): lane u holds element u in v
  auto serial = [&](uint32_t bits, float v, int cnt) {
    float el[xsum::kSeg];
#pragma unroll
    for (int u = 0; u < xsum::kSeg; u++) el[u] = __shfl_sync(0xffffffffu, v, u);
    float s = xsum::u2f(bits);
#pragma unroll
    for (int u = 0; u < xsum::kSeg; u++) s = (u < cnt) ? s + el[u] : s;
    return xsum::f2u(s);
  };
  fetch(0); fetch(1); fetch(2);
  float rawn[kXsRawPf];                                           // elements of the next batch's first serial slots
  uint32_t sern = 0u;
#pragma unroll
  for (int i = 0; i < kXsRawPf; i++) rawn[i] = 0.f;
  uint32_t bits = 0u;                                             // s = +0.f
  for (int b = 0; b < nb; b++) {
    fetch(b + 3);
    asm volatile("cp.async.wait_group 2;" ::: "memory");          // batches <= b + 1 have landed
    __syncwarp();
    const uint32_t* sb = ring[b % kXsRing];
    uint32_t w[8];                                                // own table (event slots are read from the ring)
    {
      const uint4 a0 = *reinterpret_cast<const uint4*>(sb + lane * xsum::kSlotWords);
      const uint4 a1 = *reinterpret_cast<const uint4*>(sb + lane * xsum::kSlotWords + 4);
      w[0] = a0.x; w[1] = a0.y; w[2] = a0.z; w[3] = a0.w; w[4] = a1.x; w[5] = a1.y; w[6] = a1.z; w[7] = a1.w;
    }
    const int ns = min(32, nseg - b * 32);                        // summaries in this batch
    const bool plain = w[0] == xsum::kOne;
    const uint32_t special = __ballot_sync(0xffffffffu, !plain || lane >= ns);
    float raw[kXsRawPf];
#pragma unroll
    for (int i = 0; i < kXsRawPf; i++) raw[i] = rawn[i];
    const uint32_t sermask = (b == 0) ? 0u : sern;                // batch 0 has no prefetch: its serial slots load on demand
    if (b + 1 < nb) {                                             // elements of the next batch's serial slots, used one batch later
      sern = __ballot_sync(0xffffffffu, (b + 1) * 32 + lane < nseg &&
                                            ring[(b + 1) % kXsRing][lane * xsum::kSlotWords] == xsum::kSerial);
      uint32_t m = sern;
#pragma unroll
      for (int i = 0; i < kXsRawPf; i++) {
        rawn[i] = 0.f;
        if (m) {
          const int k = __ffs(m) - 1;
          m &= m - 1u;
          rawn[i] = element((unsigned)((b + 1) * 32 + k));
        }
      }
    }
    auto redo = [&](int k) {                                      // slot k of this batch as 32 float additions
      const unsigned g = (unsigned)(b * 32 + k);
      const unsigned K0 = g * xsum::kSeg;
      const int cnt = K0 >= N ? 0 : (N - K0 >= (unsigned)xsum::kSeg ? xsum::kSeg : (int)(N - K0));
      const int idx = __popc(sermask & ((1u << k) - 1u));
      float v;
      if (((sermask >> k) & 1u) && idx < kXsRawPf) {
        v = raw[0];
#pragma unroll
        for (int i = 1; i < kXsRawPf; i++) v = (idx == i) ? raw[i] : v;
      } else {
        v = element(g);
      }
      bits = serial(bits, v, cnt);
    };
    int cur = 0;
    bool fresh = true;                                            // cur is the first slot of its stretch
    while (cur < ns) {
      const uint32_t rest = special >> cur;
      const int f = rest ? cur + __ffs(rest) - 1 : 32;            // next slot that is not a plain table (or the batch end)
      if (f > cur) {
        bool crossed = false;
        if (fresh) {                                              // the whole stretch in one table
          const uint4 t0 = *reinterpret_cast<const uint4*>(sb + (f - 1) * xsum::kSlotWords + 8);
          const uint4 t1 = *reinterpret_cast<const uint4*>(sb + (f - 1) * xsum::kSlotWords + 12);
          bool ok = true;
          const uint32_t nb_ = xsum::apply_table(bits, t0.x, (int32_t)t0.y, (int32_t)t0.z, (int32_t)t0.w, (int32_t)t1.x,
                                                 (int32_t)t1.y, (int32_t)t1.z, ok);
          if (ok) { bits = nb_; crossed = true; }
        }
        if (!crossed) {                                           // slot by slot
          uint32_t acc = bits, mine = bits;
          for (int k0 = cur; k0 < f; k0++) {
            const uint2 d = *reinterpret_cast<const uint2*>(sb + k0 * xsum::kSlotWords + 2);   // (D0, D1); 0, 0 for an empty table
            mine = (lane == k0) ? acc : mine;
            acc += (acc & 1u) ? d.y : d.x;
          }
          bool ok = true;
          if (lane >= cur && lane < f)
            xsum::apply_table(mine, w[1], (int32_t)w[2], (int32_t)w[3], (int32_t)w[4], (int32_t)w[5], (int32_t)w[6],
                              (int32_t)w[7], ok);
          const uint32_t bad = __ballot_sync(0xffffffffu, !ok);
          if (bad) {
            const int g = __ffs(bad) - 1;
            bits = __shfl_sync(0xffffffffu, mine, g);
            redo(g);
            cur = g + 1;
            fresh = false;
            continue;
          }
          bits = acc;
        }
      }
      if (f < ns) {                                               // the event slot
        if (!xsum::apply_segment(bits, sb + f * xsum::kSlotWords)) redo(f);
      }
      cur = f + 1;
      fresh = true;
    }
    __syncwarp();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (lane == 0) q.sc.psum[e] = xsum::u2f(bits);
}
