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

// serial index K (i-major over the interior) -> offset in the pitched array
struct XsCursor {
  int i, j, len, P;
  __device__ __forceinline__ XsCursor(long long K, int len_, int P_) : len(len_), P(P_) {
    i = 1 + (int)(K / len_); j = 1 + (int)(K % len_);
  }
  __device__ __forceinline__ size_t off() const { return (size_t)i * P + j; }
  __device__ __forceinline__ void advance(int d) { j += d; while (j > len) { j -= len; i++; } }
};

__global__ void __launch_bounds__(kXsThreads)
k_xsum_totals(const __grid_constant__ SolverParams q) {
  __shared__ double wsum[kXsThreads / 32];
  const int c = blockIdx.x, e = blockIdx.y, t = threadIdx.x;
  const int len = q.m - 2;
  const long long N = (long long)(q.n - 2) * len, base = (long long)c * kXsChunk;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  double s = 0.0;
  XsCursor cur(base + t, len, q.P);
#pragma unroll 4
  for (int u = 0; u < xsum::kSeg; u++) {
    if (base + t + (long long)u * kXsThreads < N) s += (double)p[cur.off()];
    cur.advance(kXsThreads);
  }
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

__global__ void __launch_bounds__(kXsThreads)
k_xsum_tables(const __grid_constant__ SolverParams q) {
  __shared__ float buf[kXsThreads * kXsPad];
  __shared__ double wsum[kXsThreads / 32];
  __shared__ double base_pred;
  const int c = blockIdx.x, e = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int len = q.m - 2;
  const long long N = (long long)(q.n - 2) * len, base = (long long)c * kXsChunk;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  {
    XsCursor cur(base + t, len, q.P);
#pragma unroll 4
    for (int u = 0; u < xsum::kSeg; u++) {
      const int kl = t + u * kXsThreads;                          // local index: segment kl / 32, element kl % 32
      buf[(kl >> 5) * kXsPad + (kl & 31)] = (base + kl < N) ? p[cur.off()] : 0.f;
      cur.advance(kXsThreads);
    }
  }
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
  if (cnt > 0) {
    uint32_t slot[xsum::kSlotWords];
    xsum::build_segment([&](int k) { return seg[k]; }, cnt, pred, slot);
    uint4* out = reinterpret_cast<uint4*>(q.xs_slots + ((size_t)e * q.xs_nseg + g) * xsum::kSlotWords);
#pragma unroll
    for (int w = 0; w < 4; w++) out[w] = make_uint4(slot[4 * w], slot[4 * w + 1], slot[4 * w + 2], slot[4 * w + 3]);
  }
}

// Serial pass.  One warp per environment; lane L holds summary 32 b + L of batch b in registers.  Summaries of
// type kOne only need  bits += (bits & 1) ? D1 : D0  on the critical path (about two dependent integer
// operations), so the warp first runs that chain over a stretch of slots -- every lane keeps the accumulator
// its own slot started from -- and then all lanes check their slot's validity condition at once.  Anything else
// (a split summary, a serial one, a summary whose condition fails) is an "event" handled on its own before the
// chain resumes behind it.
__global__ void __launch_bounds__(32)
k_xsum_chain(const __grid_constant__ SolverParams q) {
  __shared__ __align__(16) uint2 sD[2][40];                       // (D0, D1 - D0) of the batch, broadcast to the chain
  const int e = blockIdx.x, lane = threadIdx.x;
  const int len = q.m - 2, P = q.P;
  const long long N = (long long)(q.n - 2) * len;
  const int nseg = q.xs_nseg;
  const float* p = q.lev[0].x + (size_t)e * q.stride;
  const uint4* slots = reinterpret_cast<const uint4*>(q.xs_slots + (size_t)e * nseg * xsum::kSlotWords);
  uint4 nxt[4];
  auto fetch = [&](int b) {
    const int g = b * 32 + lane;
    if (g < nseg) {
#pragma unroll
      for (int w = 0; w < 4; w++) nxt[w] = slots[(size_t)g * 4 + w];
    } else {                                                      // past the end: an empty table
      nxt[0] = make_uint4(xsum::kOne, xsum::kAnyKey, 0u, 0u);
      nxt[1] = nxt[2] = nxt[3] = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  // genuine serial additions of segment g from accumulator `bits` (warp-uniform)
  auto serial = [&](uint32_t bits, long long g) {
    const long long K0 = g * xsum::kSeg, left = N - K0;
    const int cnt = left >= xsum::kSeg ? xsum::kSeg : (left > 0 ? (int)left : 0);
    float v = 0.f;
    if (lane < cnt) {
      const long long K = K0 + lane;
      v = p[(size_t)(1 + (int)(K / len)) * P + 1 + (int)(K % len)];
    }
    float s = xsum::u2f(bits);
    for (int u = 0; u < cnt; u++) s += __shfl_sync(0xffffffffu, v, u);
    return xsum::f2u(s);
  };
  const int nb = (nseg + 31) / 32;
  fetch(0);
  uint32_t bits = 0u;                                             // s = +0.f
  for (int b = 0; b < nb; b++) {
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 4; k++) { w[4 * k] = nxt[k].x; w[4 * k + 1] = nxt[k].y; w[4 * k + 2] = nxt[k].z; w[4 * k + 3] = nxt[k].w; }
    if (b + 1 < nb) fetch(b + 1);                                 // in flight during the chain
    uint2* sd = sD[b & 1];
    const bool plain = w[0] == xsum::kOne;
    sd[lane] = (plain && w[1] != xsum::kAnyKey) ? make_uint2(w[2], w[3] - w[2]) : make_uint2(0u, 0u);
    const uint32_t special = __ballot_sync(0xffffffffu, !plain);
    __syncwarp();
    int cur = 0;
    while (cur < 32) {
      const uint32_t rest = special >> cur;
      const int f = rest ? cur + __ffs(rest) - 1 : 32;            // next slot that is not a plain table
      uint32_t acc = bits, mine = bits;
      for (int k0 = cur; k0 < f; k0 += 4) {
        const uint2 d0 = sd[k0], d1 = sd[k0 + 1], d2 = sd[k0 + 2], d3 = sd[k0 + 3];
        mine = (lane == k0) ? acc : mine;
        acc = (acc + d0.x) + (acc & 1u) * d0.y;
        if (k0 + 1 < f) { mine = (lane == k0 + 1) ? acc : mine; acc = (acc + d1.x) + (acc & 1u) * d1.y; }
        if (k0 + 2 < f) { mine = (lane == k0 + 2) ? acc : mine; acc = (acc + d2.x) + (acc & 1u) * d2.y; }
        if (k0 + 3 < f) { mine = (lane == k0 + 3) ? acc : mine; acc = (acc + d3.x) + (acc & 1u) * d3.y; }
      }
      // every lane of the stretch checks that its table really applied to the accumulator it started from
      bool ok = true;
      if (lane >= cur && lane < f) xsum::apply_table(mine, w[1], (int32_t)w[2], (int32_t)w[3], (int32_t)w[4], (int32_t)w[5],
                                                     (int32_t)w[6], (int32_t)w[7], ok);
      const uint32_t bad = __ballot_sync(0xffffffffu, !ok);
      int ev;                                                     // slot to handle on its own (32 = none)
      if (bad) { ev = __ffs(bad) - 1; bits = __shfl_sync(0xffffffffu, mine, ev); }
      else { ev = f; bits = acc; }
      if (ev < 32) {
        uint32_t sw[16];
#pragma unroll
        for (int k = 0; k < 16; k++) sw[k] = __shfl_sync(0xffffffffu, w[k], ev);
        if (bad || !xsum::apply_segment(bits, sw)) bits = serial(bits, (long long)b * 32 + ev);
      }
      cur = ev + 1;
    }
    __syncwarp();
  }
  if (lane == 0) q.sc.psum[e] = xsum::u2f(bits);
}
