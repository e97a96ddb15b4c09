// smooth_strip.cuh -- exact lexicographic Gauss-Seidel smoothing (MG.smooth, MG.pde:79-89) as a
// register-pipelined strip sweep.
//
// The serial reference updates d[i][j] in place with i outer, j inner, so cell (i,j) sees NEW values at
// (i-1,j), (i,j-1) and OLD values at (i+1,j), (i,j+1).  Mapping used here, per level:
//   * a warp owns a strip of 32 consecutive columns j (contiguous in memory); lane L = global column
//     index - 1 carries column j for every row;
//   * time is skewed by one step per column: at global step t, pipeline stage g of lane L works on row
//         i = t - L - 2g + 1
//     stage 0 forms d = r*inv (MG.pde:80); stages 1..4 are the four Gauss-Seidel sweeps, each trailing
//     the previous one by two rows, which is exactly the distance at which sweep s finds sweep s-1's
//     values at (i+1,j), (i,j+1) and its own new values at (i-1,j), (i,j-1);
//   * every operand therefore comes from a register of the same lane or of a neighbouring lane one
//     step earlier: W = own previous result of the same stage, S = shfl_up of it, E = own previous result
//     of the stage before, N = shfl_down of that.  Strip boundaries continue the skew (strip k runs
//     32 steps behind strip k-1), so lane 31 / lane 0 of adjacent warps exchange the same registers
//     through a parity-double-buffered shared-memory mailbox with one __syncthreads per step;
//   * out-of-domain operands are ghosts of d, which during the sweeps hold r_ghost*inv_ghost; their
//     products with the boundary face coefficients are +-0 on every level (level 0: r_ghost = 0; coarse
//     levels: boundary coefficients = 0, MG.pde:120).  The pre-skewed coefficient tables hold zeros for
//     every (step, lane) that is not an interior cell, so such a stage evaluates to +-0 by itself and no
//     activity predicate is needed in the sweeps;
//   * static coefficients are stored pre-skewed ([strip][step][lane], built on the host, L2-resident);
//     cp.async lands each set in a small shared-memory ring kPrefetch steps ahead (a register-target
//     LDG prefetch that deep serialises on the six hardware scoreboards), from where it enters a
//     register queue and is reused by the four sweeps;
//     r is staged through a per-warp shared-memory ring filled by cp.async with coalesced row reads and
//     then travels with the coefficients in the register queue; finished rows of d leave through a
//     second ring and are written as coalesced rows.
// Arithmetic per update keeps the reference order:
//     d = -(dW*lxW + dE*lxE + dS*lyS + dN*lyN - r) * inv      (the minus sign is folded into ninv = -inv,
//                                                               which is exact: (-a)*b == a*(-b) bitwise)
#pragma once
#include <cuda_runtime.h>

#include "solver.h"

namespace rlfc {

constexpr int kRing = 40;          // ring depth in rows (>= 32 lanes of skew + 8 rows of stages)
constexpr int kPrefetch = 6;       // cp.async groups in flight
constexpr int kCoefRing = 8;       // shared-memory landing ring of the coefficient sets (power of two > kPrefetch)
constexpr int kLook = 1;           // coefficient sets enter the register queue kLook steps before their first use
constexpr int kQueue = 12;         // register queue depth = unroll factor (coefficient ages kLook + 0..10, r ages 0..10)
constexpr int kSkewPad = 16;       // front padding of the pre-skewed tables: entry 0 is tau = -kSkewPad
constexpr int kSkewTail = 64;      // back padding: tables cover tau < ni + kSkewTail
constexpr int kWarm = kQueue;      // a strip starts executing at tau = -kWarm (even, multiple of kQueue)
constexpr unsigned kRowBytes = 32 * 4;
constexpr unsigned kRingBytes = kRing * kRowBytes;
__device__ __forceinline__ unsigned ring_next(unsigned o) { return (o == kRingBytes - kRowBytes) ? 0u : o + kRowBytes; }
__host__ __device__ constexpr unsigned ring_slot(int row) { return (unsigned)(((row % kRing) + kRing) % kRing) * kRowBytes; }

struct __align__(16) StripShared { // per strip (warp)
  float4 cf[kCoefRing][32];        // coefficient sets (lxW, lxE, lyS, lyN) landing here by cp.async; slot = entry & 7
  float2 cn[kCoefRing][32];        // (ninv, diag)
  float in[kRing][32];             // r rows, written by cp.async (lane l owns column l); slot = row mod kRing
  float out[kRing][32];            // finished rows of d / x / p; slot = (row + 8) mod kRing, (row + 10) for XMODE 3
  float out2[kRing][32];           // XMODE 3: finished rows of the new residual
};

struct __align__(16) StripMail {   // slot k+1 belongs to strip k; slots 0 and nstrips+1 stay zero (domain edges)
  float4 hi[2];                    // lane 31's stages 1..4, by step parity  -> next strip's lane 0
  float4 lo[2];                    // lane 0's stages 0..3                   -> previous strip's lane 31
  float hi5[2];                    // XMODE 3: lane 31's final d three steps back (row of the increment stage)
  float lo5[2];                    // XMODE 3: lane 0's stage 4
  float pad[4];
};

// ---- explicit shared-memory access on 32-bit shared addresses ----
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void sts32_if(bool p, unsigned a, float v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f32 [%0], %1; }" ::"r"(a), "f"(v), "r"((unsigned)p));
}
__device__ __forceinline__ float2 lds64(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128_if(bool p, unsigned a, float x, float y, float z, float w) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.shared.v4.f32 [%0], {%1,%2,%3,%4}; }" ::"r"(a), "f"(x), "f"(y),
               "f"(z), "f"(w), "r"((unsigned)p));
}
__device__ __forceinline__ void cp_async4_if(bool p, unsigned smem, const void* gmem) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q cp.async.ca.shared.global [%0], [%1], 4; }" ::"r"(smem), "l"(gmem),
               "r"((unsigned)p));
}
__device__ __forceinline__ void cp_async4(unsigned smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(unsigned smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(unsigned smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct StripState {
  float prev[5];                   // result of stage g at the previous step
  float h2, h3;                    // XMODE 3: stage-4 results two and three steps back
  float4 qa[kQueue];               // (lxW, lxE, lyS, lyN), slot = step mod kQueue of the step that loaded it
  float2 qn[kQueue];               // (ninv, diag)
  float qr[kQueue];                // r of stage 0's row
  const float4* pa;                // running table pointers: entry tau + kLook + kPrefetch of this strip / lane
  const float2* pn;
  const float* pr;                 // r row to prefetch next (row tau + 1 + kPrefetch, this lane's column)
  float* pd;                       // row of d / x / p to write out next
  float* pr_out;                   // XMODE 3: row of r to write out next
  const float* px;                 // XMODE 2/3: x (p) row prefetched into the out ring (row tau-1 / tau-3)
  unsigned o_row;                  // ring offset of stage 0's row tau - lane + 1 (= slot of the row the last stage emits)
  unsigned o_pf, o_w;              // ring offsets of the row being prefetched / written out
  unsigned a_in, a_out, a_out2;    // shared addresses of this lane's column in the rings
  unsigned a_cf, a_cn;             // shared addresses of this lane's column in the coefficient landing ring
  unsigned a_mh, a_ml, a_hi, a_lo; // mailboxes: previous strip's hi, next strip's lo, own hi, own lo (parity 0)
  unsigned a_gtop, a_gleft;        // XMODE 3: this column's slot in the top-row buffer (bottom = +4*mj), base of the
                                   // left-column buffer (right = +4*ni)
  double rr;                       // XMODE 3: this lane's share of r.r
};

struct StripEdges {                // XMODE 3: per-lane flags / sizes of the increment stage
  int ni, mj;
  bool jfirst, jlast;              // this lane's column is 1 / mj
};

// one global step of one strip; PH = t mod kQueue resolves the register queue slots (and the step parity) at
// compile time
// XMODE selects what happens with a finished row of d:
//   0: d_out = d                      (plain smoother output)
//   1: x = 0 + d                      (coarsest level: x starts at 0, MG.pde:56,95)
//   2: x = x + d                      (x.plusEq(d), MG.pde:95); x rows are prefetched into the out ring and the sum is
//                                     formed by stage 4, so the write-out never waits on a global load
//   3: level-0 smooth(4) complete (MG.pde:90-97): a fifth stage two rows behind the last sweep applies
//      d.setBC (clamped neighbours), x += d, r -= A d, accumulates r.r, and keeps the boundary values
//      of d for the ghost cells of x
template <int PH, int XMODE>
__device__ __forceinline__ void strip_step(StripState& s, int tau, int lane, unsigned ni_eff, int lag, int P,
                                           const StripEdges& ed) {
  constexpr unsigned par = PH & 1, rd = (par ^ 1) * 16, wr = par * 16, rd4 = (par ^ 1) * 4, wr4 = par * 4;
  // ---- neighbour operands as of the end of the previous step ----
  float up[5], dn[5], up5 = 0.f;
#pragma unroll
  for (int g = 1; g <= 4; g++) up[g] = __shfl_up_sync(0xffffffffu, s.prev[g], 1);
#pragma unroll
  for (int g = 0; g <= (XMODE == 3 ? 4 : 3); g++) dn[g] = __shfl_down_sync(0xffffffffu, s.prev[g], 1);
  if (XMODE == 3) up5 = __shfl_up_sync(0xffffffffu, s.h3, 1);
  {
    const float4 mh = lds128(s.a_mh + rd), ml = lds128(s.a_ml + rd);
    const bool l0 = lane == 0, l31 = lane == 31;
    up[1] = l0 ? mh.x : up[1]; up[2] = l0 ? mh.y : up[2]; up[3] = l0 ? mh.z : up[3]; up[4] = l0 ? mh.w : up[4];
    dn[0] = l31 ? ml.x : dn[0]; dn[1] = l31 ? ml.y : dn[1]; dn[2] = l31 ? ml.z : dn[2]; dn[3] = l31 ? ml.w : dn[3];
    if (XMODE == 3) {
      const float m5h = lds32(s.a_mh + 64 + rd4), m5l = lds32(s.a_ml + 40 + rd4);   // hi5 at +64, lo5 at +72 from hi[0]
      up5 = l0 ? m5h : up5;
      dn[4] = l31 ? m5l : dn[4];
    }
  }
  // ---- prefetch (cp.async) r row tau+1+kPrefetch and coefficient entry tau+kLook+kPrefetch; then move
  //      coefficient entry tau+kLook from its landing slot into the register queue ----
  {
    const unsigned cslot = (unsigned)(tau + kLook + kPrefetch) & (kCoefRing - 1);
    cp_async4_if((unsigned)(tau + kPrefetch) < ni_eff, s.a_in + s.o_pf, s.pr);
    if (XMODE == 2) {
      // x of row tau-1 lands in the out-ring slot that stage 4 will fill for that row (slot (row + 8) mod kRing is
      // the slot of r row tau+7); lane 0 reaches the row at step row + 7 = tau + kPrefetch, the group's deadline
      cp_async4_if((unsigned)(tau - 2) < ni_eff, s.a_out + s.o_pf, s.px);
      s.px += P;
    }
    if (XMODE == 3) {
      // p of row tau-3: slot (row + 10) mod kRing; lane 0's increment stage reaches it at step row + 9 = tau + kPrefetch
      cp_async4_if((unsigned)(tau - 4) < ni_eff, s.a_out + s.o_pf, s.px);
      s.px += P;
    }
    cp_async16(s.a_cf + cslot * 512u, s.pa);
    cp_async8(s.a_cn + cslot * 256u, s.pn);
    cp_async_commit();
    s.o_pf = ring_next(s.o_pf);
    s.pr += P;
    s.pa += 32;
    s.pn += 32;
    cp_async_wait<kPrefetch>();
    const unsigned rslot = (unsigned)(tau + kLook) & (kCoefRing - 1);
    s.qa[PH] = lds128(s.a_cf + rslot * 512u);
    s.qn[PH] = lds64(s.a_cn + rslot * 256u);
  }
  float res[5];
  // ---- stage 0: d = r * inv on row i0 = tau - lane + 1 ----
  {
    float rv = lds32(s.a_in + s.o_row);
    rv = ((unsigned)(tau - lane) < ni_eff) ? rv : 0.f;
    s.qr[PH] = rv;
    res[0] = rv * (-s.qn[(PH - kLook + kQueue) % kQueue].x);
  }
  // ---- stages 1..4: the Gauss-Seidel sweeps (zero coefficients make non-cells evaluate to +-0) ----
#pragma unroll
  for (int g = 1; g <= 4; g++) {
    const float4 c = s.qa[(PH - 2 * g - kLook + 2 * kQueue) % kQueue];
    const float ninv = s.qn[(PH - 2 * g - kLook + 2 * kQueue) % kQueue].x;
    const float rv = s.qr[(PH - 2 * g + kQueue) % kQueue];
    res[g] = (s.prev[g] * c.x + s.prev[g - 1] * c.y + up[g] * c.z + dn[g - 1] * c.w - rv) * ninv;
  }
  if (XMODE == 2) sts32(s.a_out + s.o_row, lds32(s.a_out + s.o_row) + res[4]);   // x.plusEq(d): x + d
  else if (XMODE != 3) sts32(s.a_out + s.o_row, res[4]);   // row tau - lane - 7 lives in slot (row + 8) mod kRing
  if (XMODE == 3) {
    // ---- stage 5 on row i5 = tau - lane - 9: d.setBC (ghost = adjacent interior), x += d, r -= A d ----
    const int i5 = tau - lane - 9;
    const bool valid = (unsigned)(i5 - 1) < ni_eff;
    const float dC = s.h2;
    const float dW = (i5 == 1) ? dC : s.h3;
    const float dE = (i5 == ed.ni) ? dC : s.prev[4];
    const float dS = ed.jfirst ? dC : up5;
    const float dN = ed.jlast ? dC : dn[4];
    const float4 c = s.qa[(PH - 10 - kLook + 2 * kQueue) % kQueue];
    const float diag = s.qn[(PH - 10 - kLook + 2 * kQueue) % kQueue].y;
    const float rv = s.qr[(PH - 10 + kQueue) % kQueue];
    const float Ad = dC * diag + dW * c.x + dE * c.y + dS * c.z + dN * c.w;       // PoissonMatrix.pde:56-61
    const float rN = rv - Ad;
    sts32(s.a_out + s.o_row, lds32(s.a_out + s.o_row) + dC);                     // x + d, row i5 in slot (row + 10)
    sts32(s.a_out2 + s.o_row, rN);
    // boundary values of d feed the ghost cells of x after the sweep (x.plusEq(d) runs over all cells)
    sts32_if(valid && i5 == 1, s.a_gtop, dC);
    sts32_if(valid && i5 == ed.ni, s.a_gtop + 4u * ed.mj, dC);
    sts32_if(valid && ed.jfirst, s.a_gleft + 4u * (unsigned)(i5 - 1), dC);
    sts32_if(valid && ed.jlast, s.a_gleft + 4u * (unsigned)(ed.ni + i5 - 1), dC);
    s.h3 = s.h2;
    s.h2 = s.prev[4];
  }
  s.o_row = ring_next(s.o_row);
#pragma unroll
  for (int g = 0; g <= 4; g++) s.prev[g] = res[g];
  // ---- mailbox for the neighbouring strips ----
  sts128_if(lane == 31, s.a_hi + wr, res[1], res[2], res[3], res[4]);
  sts128_if(lane == 0, s.a_lo + wr, res[0], res[1], res[2], res[3]);
  if (XMODE == 3) {
    sts32_if(lane == 31, s.a_hi + 64 + wr4, s.h3);
    sts32_if(lane == 0, s.a_lo + 40 + wr4, res[4]);
  }
  // ---- write out the row every lane of the strip has finished (coalesced) ----
  {
    const int w = tau - lag;
    const float v = lds32(s.a_out + s.o_w);
    const bool ok = (unsigned)(w - 1) < ni_eff;
    if (XMODE == 3) {
      const float rN = lds32(s.a_out2 + s.o_w);
      if (ok) {
        *s.pd = v;
        *s.pr_out = rN;
        const float prod = rN * rN;                 // float product, double accumulation (Field.pde:304-307)
        s.rr += (double)prod;
      }
      s.pr_out += P;
    } else if (ok) {
      *s.pd = (XMODE == 1) ? 0.f + v : v;
    }
    s.o_w = ring_next(s.o_w);
    s.pd += P;
  }
}

// Smooth one level for one environment: d(interior) = four lexicographic GS sweeps started from r*inv.
// Called by ALL threads of the CTA (warps beyond the level's strip count only take part in the
// barriers).  r and d are this environment's row-major pitched arrays of the level (d is the x array for
// XMODE 1, 2 and 3; in XMODE 3 the new residual goes to r_out (which may alias r) and gbuf (shared, 2*mj + 2*ni
// floats) receives d on the four boundary lines: top row, bottom row, left column, right column).  Returns this
// lane's share of r.r.
template <int XMODE>
__device__ __forceinline__ double strip_smooth(const DevLevel& L, const float* __restrict__ r, float* __restrict__ d,
                                               StripShared* sh_all, StripMail* mail, float* gbuf = nullptr,
                                               float* __restrict__ r_out = nullptr) {
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
  const int ni = L.n - 2, mj = L.m - 2, P = L.P;
  const int nstrips = L.sk.nstrips;
  const bool mine = k < nstrips;
  const int j = 32 * k + lane + 1;
  const unsigned ni_eff = (mine && j <= mj) ? (unsigned)ni : 0u;
  constexpr int kStages = (XMODE == 3) ? 10 : 8;   // rows between stage 0 and the emitting stage, plus one
  const int lag = min(32, mj) + kStages;        // lane min(31,mj-1) emits row w at tau = w + min(31,mj-1) + kStages - 1
  const int tau_end = ni + lag;                 // last write-out
  // zero the mailboxes (slots 0 .. nstrips+1)
  for (int c = threadIdx.x; c < (nstrips + 2) * (int)(sizeof(StripMail) / 4); c += blockDim.x)
    reinterpret_cast<float*>(mail)[c] = 0.f;
  StripState s;
#pragma unroll
  for (int g = 0; g < 5; g++) s.prev[g] = 0.f;
  s.h2 = 0.f; s.h3 = 0.f; s.rr = 0.0;
#pragma unroll
  for (int a = 0; a < kQueue; a++) { s.qa[a] = make_float4(0.f, 0.f, 0.f, 0.f); s.qn[a] = make_float2(0.f, 0.f); s.qr[a] = 0.f; }
  StripShared& sh = sh_all[mine ? k : 0];
  StripMail* mymail = mail + 1 + (mine ? k : 0);
  StripEdges ed;
  ed.ni = ni; ed.mj = mj; ed.jfirst = (j == 1); ed.jlast = (j == mj);
  {
    const size_t base = ((size_t)(mine ? k : 0) * L.sk.Tsk + (-kWarm + kLook + kSkewPad)) * 32 + lane;
    s.pa = L.sk.A + base;           // entry -kWarm + kLook; advanced past the initial fill below
    s.pn = L.sk.nd + base;
    const int jj = mine ? min(j, L.m - 1) : 1;
    s.pr = r + (ptrdiff_t)(-kWarm + 1 + kPrefetch) * P + jj;
    s.pd = d + (ptrdiff_t)(-kWarm - lag) * P + jj;
    s.pr_out = (XMODE == 3 ? r_out : d) + (ptrdiff_t)(-kWarm - lag) * P + jj;
    s.o_row = ring_slot(-kWarm - lane + 1);
    s.o_pf = ring_slot(-kWarm + 1 + kPrefetch);
    s.o_w = ring_slot(-kWarm - lag + kStages);
    s.px = d + (ptrdiff_t)(-kWarm - (XMODE == 3 ? 3 : 1)) * P + jj;
    s.a_cf = smem_addr(&sh.cf[0][lane]);
    s.a_cn = smem_addr(&sh.cn[0][lane]);
    s.a_in = smem_addr(&sh.in[0][lane]);
    s.a_out = smem_addr(&sh.out[0][lane]);
    s.a_out2 = smem_addr(&sh.out2[0][lane]);
    s.a_mh = smem_addr(&mymail[-1].hi[0]);
    s.a_ml = smem_addr(&mymail[1].lo[0]);
    s.a_hi = smem_addr(&mymail->hi[0]);
    s.a_lo = smem_addr(&mymail->lo[0]);
    s.a_gtop = (XMODE == 3) ? smem_addr(gbuf + min(j, mj) - 1) : 0u;
    s.a_gleft = (XMODE == 3) ? smem_addr(gbuf + 2 * mj) : 0u;
    // pin the loop invariants in registers (otherwise they are rematerialised from the kernel parameters every step)
    asm volatile("" : "+r"(s.a_in), "+r"(s.a_out), "+r"(s.a_mh), "+r"(s.a_ml), "+r"(s.a_hi), "+r"(s.a_lo), "+r"(s.a_cf),
                 "+r"(s.a_cn), "+r"(s.a_out2));
  }
  unsigned ni_pin = ni_eff;
  int lag_pin = lag, P_pin = P, lane_pin = lane;
  asm volatile("" : "+r"(ni_pin), "+r"(lag_pin), "+r"(P_pin), "+r"(lane_pin));
  // initial fill: r rows 1..kPrefetch and coefficient entries -kWarm+kLook .. -kWarm+kLook+kPrefetch-1; later
  // rows / entries are issued step by step
  if (mine) {
    for (int rho = 1; rho <= kPrefetch; rho++)
      cp_async4_if((unsigned)(rho - 1) < ni_eff, s.a_in + ring_slot(rho), r + IDX(rho, j));
    for (int c = 0; c < kPrefetch; c++) {
      const unsigned cslot = (unsigned)(-kWarm + kLook + c) & (kCoefRing - 1);
      cp_async16(s.a_cf + cslot * 512u, s.pa);
      cp_async8(s.a_cn + cslot * 256u, s.pn);
      s.pa += 32;
      s.pn += 32;
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  // global steps t = -kWarm .. 32*(nstrips-1) + tau_end; strip k executes while -kWarm <= t - 32k <= tau_end
  const int t_last = 32 * (nstrips - 1) + tau_end;
  const int off = 32 * k;
  for (int t0 = -kWarm; t0 <= t_last; t0 += kQueue) {
#define RLFC_STRIP_STEP(PH_)                                                                        \
    {                                                                                               \
      const int tau = t0 + PH_ - off;                                                               \
      if (mine && tau >= -kWarm && tau <= tau_end)                                                   \
        strip_step<PH_, XMODE>(s, tau, lane_pin, ni_pin, lag_pin, P_pin, ed);                        \
      __syncthreads();                                                                              \
    }
    RLFC_STRIP_STEP(0) RLFC_STRIP_STEP(1) RLFC_STRIP_STEP(2) RLFC_STRIP_STEP(3) RLFC_STRIP_STEP(4) RLFC_STRIP_STEP(5)
    RLFC_STRIP_STEP(6) RLFC_STRIP_STEP(7) RLFC_STRIP_STEP(8) RLFC_STRIP_STEP(9) RLFC_STRIP_STEP(10) RLFC_STRIP_STEP(11)
#undef RLFC_STRIP_STEP
  }
  cp_async_wait<0>();
  __syncthreads();
  return s.rr;
}

}  // namespace rlfc
