// smooth_rows.cuh -- exact lexicographic Gauss-Seidel smoothing (MG.smooth, MG.pde:79-97) as a
// row pipeline with one warp per sweep.
//
// The serial reference updates d[i][j] in place with i outer, j inner, so cell (i,j) sees NEW values at
// (i-1,j), (i,j-1) and OLD values (previous sweep) at (i+1,j), (i,j+1).  Mapping, per level and environment
// (one CTA of kRowsWarps warps; tools/rows_schedule_model.py is a numpy model of exactly this schedule):
//   * lane L owns the C = ceil(mj/32) consecutive columns j0 = C*L+1 .. C*L+C and walks down the rows, one
//     row per step; lanes are skewed by one step, so within a step the lane's C cells form the only serial
//     chain (C dependent updates) and (i,j0-1) was finished by lane L-1 one step earlier;
//   * stage g of lane L works on row  i = t - L - 2g  at global step t:  stage 0 forms d = r*inv, stages
//     1..4 are the four sweeps (each two rows behind the previous one: exactly the distance at which it
//     finds the previous sweep's values at (i+1,j), (i,j+1)), the increment stage (row t - L - 10) applies
//     d.setBC, x += d and, on level 0, r -= A d;
//   * every stage runs in its OWN warp, so a step costs one sweep of one row (C*9 flops per lane) instead
//     of five:  warps 0..3 = sweeps 1..4,  warp 4 = residual increment + r.r (level 0),  warp 5 = stage 0,
//     warp 6 = x increment + coalesced write-out,  warp 7 = bulk-copy loader.  Stages hand rows to
//     each other through small shared-memory buffers indexed by step ([lane][C] blocks, read and written
//     with vector accesses), with one __syncthreads per step;
//   * static coefficients come from a host-built, pre-skewed table: entry tau holds, for lane L, the
//     coefficients of row tau - L ([cy(C+1) | -inv(C) | cx(C)] as float4 vectors, lane-contiguous),
//     so all lanes of a stage read the same ring slot; entries land in a 16-slot shared-memory ring by
//     cp.async kPF steps ahead.  Entries are zero for every (row, column) that is not an interior cell, so
//     such a stage evaluates to +-0 by itself; out-of-domain operands are ghosts of d whose products with
//     the boundary coefficients are +-0 on every level (level 0: r_ghost = 0; coarse: boundary coefficients
//     = 0, MG.pde:120);
//   * r (and x) rows are staged in kRowRing-deep shared-memory rings with coalesced 16-byte cp.async row
//     copies (plain rows); stage 0 picks this lane's columns of row t - L out of it and republishes them in a
//     step-indexed ring for the sweeps; the increment stages write x+d and the new residual back into the
//     plain rings, and once all lanes are past a row it is written out as a coalesced row.
// Arithmetic per update keeps the reference order:
//     d = -(dW*lxW + dE*lxE + dS*lyS + dN*lyN - r) * inv      (the minus sign is folded into ninv = -inv,
//                                                               which is exact: (-a)*b == a*(-b) bitwise)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <type_traits>

#ifndef RLFC_NO_SOLVER_H
#include "solver.h"
#endif

namespace rlfc {

constexpr int kRowsWarps = 8;       // coarse V-cycle kernel: its down / up passes use every warp; the pipeline needs seven
constexpr int kRowsThreads = 32 * kRowsWarps;
#ifndef RLFC_ROWS_BULK_WO
#define RLFC_ROWS_BULK_WO 3        // x rows leave through one shared->global bulk copy in modes >= this (2: also the coarse levels)
#endif
#ifndef RLFC_ROWS_RES_SPLIT
#define RLFC_ROWS_RES_SPLIT 1      // level 0: the residual stage as two warps (lower / upper half of every lane's columns)
#endif
constexpr int kRows0Threads = 32 * 7;   // level-0 kernel: the seven pipeline warps
// warp roles of the row pipeline (a warp's scheduler is warp id mod 4)
//   level 0 (mode 3):  {residual A, x increment} {residual B, loader} {sweeps 1+2, stage 0} {sweeps 3+4}
//   modes 1, 2:        {loader, -} {x increment, -} {sweeps 3+4, stage 0} {sweeps 1+2}      (no residual stage)
template <bool SKEW> struct rows_detail_roles {
  static constexpr int ResA = SKEW ? 0 : 4, ResB = SKEW ? 1 : 5, SweepA = SKEW ? 2 : 3, SweepB = SKEW ? 3 : 2,
                       Xinc = SKEW ? 4 : 1, Loader = SKEW ? 5 : 0, Stage0 = 6;
};
#ifndef RLFC_KPF
#define RLFC_KPF 5
#endif
constexpr int kPF = RLFC_KPF;      // bulk-copy groups (steps) in flight; at most 5 with 16 coefficient slots
constexpr int kCoefShift = 4;
constexpr int kCoefSlots = 1 << kCoefShift;     // coefficient-entry ring: entries t-10 .. t+kPF live at step t
constexpr int kRSlots = 16;        // step-indexed ring of this lane's r values (stage 0 -> sweeps, increment)
constexpr int kRowRing = 48;       // plain r / x row rings: rows t+kPF .. t-32-10 live at step t
constexpr int kStageLag = 10;      // rows between stage 0 and the residual increment stage
constexpr int kTabFront = 10;      // table entry index = tau + kTabFront (entries tau <= 0 are zero)
constexpr int kSLanes = 33;        // stage buffers carry a zero 33rd lane (right-hand domain edge)

__host__ __device__ constexpr int rows_K(int C) { return (3 * C + 1 + 3) / 4; }   // float4 vectors per lane-entry
__host__ __device__ constexpr int rows_CP(int C) { return (C + 1) & ~1; }          // lane block of the stage buffers (even)
__host__ __device__ inline int rows_table_entries(int ni, int nl) { return kTabFront + ni + nl + kStageLag + kPF + 12; }

// dynamic shared memory of one rows_smooth call (bytes): [coef ring | r ring (plain-r modes) | x ring | R ring |
// stage buffers | mbarriers]; `skewed_r` = level-0 mode, where r arrives pre-skewed and needs no row ring
__host__ __device__ inline size_t rows_smem_bytes(int C, int P, bool skewed_r) {
  return (size_t)kCoefSlots * rows_K(C) * 32 * 16 + (skewed_r ? 1 : 2) * (size_t)kRowRing * P * 4 +
         (size_t)(kRSlots * 32 + 5 * 2 * kSLanes) * rows_CP(C) * 4 + 16 + kCoefSlots * 8;
}
// skewed residual array of one environment (level 0): entry tau = i + l holds row i of lane l's columns,
// [tau][32][CP] floats; entries the smoother touches: 1 .. ni + nl + kStageLag + kPF + 2
__host__ __device__ inline size_t rows_skew_floats(int C, int ni, int nl) {
  return (size_t)(ni + nl + kStageLag + kPF + 14) * 32 * rows_CP(C);
}

namespace rows_detail {

__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- bulk async copies (cp.async.bulk, SASS UBLKCP) completing on a shared-memory mbarrier ----
__device__ __forceinline__ unsigned sm_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(sm_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(sm_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
#ifdef RLFC_DEBUG_MBAR
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > (1ll << 28)) {
      printf("mbar timeout: block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, sm_addr(bar), parity);
      __trap();
    }
  }
#else
  while (!mbar_try(bar, parity)) {}
#endif
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_addr(smem)),
               "l"(gmem), "r"(bytes), "r"(sm_addr(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global bulk copy (bulk async-group completion): the write-out of a finished row by ONE thread
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(sm_addr(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// floats [A, B) of a lane-entry (float4 vectors, lane-contiguous: vector k of lane l sits at ent[k*32 + l])
template <int A, int B>
__device__ __forceinline__ void ld_entry(const float4* ent, float (&out)[B - A]) {
  constexpr int k0 = A / 4, k1 = (B + 3) / 4;
#pragma unroll
  for (int k = k0; k < k1; k++) {
    const float4 v = ent[k * 32];
    const float tmp[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int f = 4 * k + q;
      if (f >= A && f < B) out[f - A] = tmp[q];
    }
  }
}

// a lane block of C floats (padded to CP = even) as float2 vector accesses
template <int C>
__device__ __forceinline__ void ld_block(const float* p, float (&out)[C]) {
#pragma unroll
  for (int k = 0; k < C / 2; k++) {
    const float2 v = reinterpret_cast<const float2*>(p)[k];
    out[2 * k] = v.x; out[2 * k + 1] = v.y;
  }
  if (C & 1) out[C - 1] = p[C - 1];
}
template <int C>
__device__ __forceinline__ void st_block(float* p, const float (&in)[C]) {
#pragma unroll
  for (int k = 0; k < C / 2; k++) reinterpret_cast<float2*>(p)[k] = make_float2(in[2 * k], in[2 * k + 1]);
  if (C & 1) p[C - 1] = in[C - 1];
}

// The per-step barrier of the warp-specialised pipeline: every warp arrives from its OWN role loop (different call
// sites), which is what a NAMED barrier with an explicit thread count is for (bar.sync id, count; __syncthreads in
// role-divergent code is formally undefined and is what compute-sanitizer's synccheck reports).
template <int NT>
__device__ __forceinline__ void step_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

__device__ __forceinline__ int ring_inc(int s) { return (s + 1 == kRowRing) ? 0 : s + 1; }
__host__ __device__ constexpr int ring_mod(int row) { return ((row % kRowRing) + kRowRing) % kRowRing; }

}  // namespace rows_detail

// XMODE selects what the increment stages do with a finished row of d:
//   1: x = 0 + d                      (coarsest level: x starts at 0, MG.pde:56,95)
//   2: x = x + d                      (x.plusEq(d), MG.pde:95)
//   3: level-0 smooth(4) complete (MG.pde:90-97): d.setBC (clamped neighbours), x += d, r -= A d, r.r accumulated,
//      and the boundary values of d kept in gbuf (top row, bottom row, left column, right column: 2*mj + 2*ni
//      floats) for the ghost cells of x.  In this mode `r` is the environment's SKEWED residual array
//      (rows_skew_floats; written by k_mg_up0): entry tau is one contiguous block that is bulk-copied straight into
//      the step-indexed ring, and the new residual is written back in place, skewed, as coalesced stores.
// Warp roles (seven warps, ids in rows_detail_roles; an eighth warp of the coarse kernel idles at the step barrier): sweeps 1+2, sweeps 3+4, the residual increment of the lower /
// upper half of every lane's columns (mode 3; idle otherwise), x increment, loader, stage 0.  The pipeline is bound by the
// SM's shared-memory pipe (ncu: 71-76 % of its wavefront peak with one sweep per warp, profiles/r02_rows_balance.md), so a
// warp runs TWO consecutive sweeps: the second one works two rows behind the first, i.e. on the row the first one
// finished two steps earlier, so its coefficients are the first one's of two steps ago (kept in registers) and its
// operands from the previous sweep are the first one's last two results (registers, one shuffle for the neighbouring
// lane's column): no shared-memory traffic at all for every second sweep.
// Called by ALL threads of the CTA.  x (and r in modes 1, 2) are this environment's row-major pitched
// arrays of the level.  Returns this thread's share of r.r (XMODE 3; non-zero only in the residual warp).
template <int C, int XMODE>
__device__ __forceinline__ double rows_smooth(const DevLevel& L, float* __restrict__ r, float* __restrict__ x,
                                              unsigned char* smem_raw, float* gbuf) {
  using namespace rows_detail;
  constexpr int K = rows_K(C), CP = rows_CP(C);
  // float offsets inside a lane-entry
  constexpr int F_CY = 0, F_NINV = C + 1, F_CX = 2 * C + 1, F_END = 3 * C + 1;
  constexpr bool SKEW = (XMODE == 3);
  constexpr int NT = SKEW ? kRows0Threads : kRowsThreads;
  using Roles = rows_detail_roles<SKEW>;
  constexpr int kWResA = Roles::ResA, kWResB = Roles::ResB, kWSweepA = Roles::SweepA, kWSweepB = Roles::SweepB,
                kWXinc = Roles::Xinc, kWLoader = Roles::Loader, kWStage0 = Roles::Stage0;
  constexpr bool BULK_WO = XMODE >= RLFC_ROWS_BULK_WO;   // finished x rows: one bulk copy by the loader / per-lane stores by warp 6
  constexpr int ES = K * 32;                       // float4s per entry
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ni = L.n - 2, mj = L.m - 2, P = L.P;
  const int nl = (mj + C - 1) / C;                 // lanes that own at least one column
  const int lag = nl + kStageLag;                  // row w is complete after step w + nl - 1 + kStageLag
  const int t_end = (ni + lag + 5) / 6 * 6;        // last step (a multiple of 6: the sweep warps' unrolled period; the last
                                                   // write-out is at step ni + lag; later steps only see zero table entries)
  float4* coef = reinterpret_cast<float4*>(smem_raw);
  float* rring = reinterpret_cast<float*>(smem_raw + (size_t)kCoefSlots * ES * 16);   // plain-r modes only
  float* xring = rring + (SKEW ? 0 : (size_t)kRowRing * P);
  float* R = xring + (size_t)kRowRing * P;         // [kRSlots][32][CP]
  float* S = R + (size_t)kRSlots * 32 * CP;        // [stage 0..4][parity][kSLanes][CP]
  const float4* __restrict__ tab = L.rt.T + (size_t)(blockIdx.x % (unsigned)L.rt.copies) * ((size_t)L.rt.entries * ES);
  const int j0 = C * lane + 1;                     // first column of this lane

  // zero the coefficient ring (entries tau <= 0), the R ring and the stage buffers (incl. the 33rd lane)
  for (int k = threadIdx.x; k < kCoefSlots * ES; k += NT) coef[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = threadIdx.x; k < (kRSlots * 32 + 5 * 2 * kSLanes) * CP; k += NT) R[k] = 0.f;
  __syncthreads();

  // loader (one lane of warp 5): row q of r (and x) and table entry q as bulk copies completing on bars[(q-1) & 15]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(S + (size_t)5 * 2 * kSLanes * CP) + 15) & ~(uintptr_t)15);
  if (threadIdx.x == 0) {
    for (int k = 0; k < kCoefSlots; k++) mbar_init(bars + k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned row_bytes = (unsigned)P * 4u, ent_bytes = (unsigned)ES * 16u, rsk_bytes = 32u * CP * 4u;
  auto issue = [&](int q) {
    unsigned long long* bar = bars + ((q - 1) & (kCoefSlots - 1));    // q >= 1: use k = (q-1)/16 of this barrier
    const bool row = q >= 1 && q <= ni;
    mbar_expect_tx(bar, ent_bytes + (SKEW ? rsk_bytes : 0u) + (row ? ((XMODE != 1 ? 1u : 0u) + (SKEW ? 0u : 1u)) * row_bytes : 0u));
    if (row) {
      const int slot = ring_mod(q);
      if (!SKEW) bulk_g2s(rring + (size_t)slot * P, r + (size_t)q * P, row_bytes, bar);
      if (XMODE != 1) bulk_g2s(xring + (size_t)slot * P, x + (size_t)q * P, row_bytes, bar);
    }
    if (SKEW) bulk_g2s(R + (size_t)(q & (kRSlots - 1)) * 32 * CP, r + (size_t)q * 32 * CP, rsk_bytes, bar);
    bulk_g2s(coef + (size_t)(q & (kCoefSlots - 1)) * ES, tab + (size_t)(q + kTabFront) * ES, ent_bytes, bar);
  };
  if (warp == kWLoader) {
    if (lane == 0) {
      fence_proxy_async();
      for (int q = 1; q <= kPF; q++) issue(q);
    }
    mbar_wait(bars + 0, 0);
  }
  __syncthreads();

#ifdef RLFC_ROWS_TIMING        // per-warp busy time between two step barriers (tools: which stage does a step wait for?)
  long long busy_ = 0, tk_ = clock64();
#define RLFC_STEP_SYNC() do { busy_ += clock64() - tk_; rows_detail::step_barrier<NT>(); tk_ = clock64(); } while (0)
#else
#define RLFC_STEP_SYNC() rows_detail::step_barrier<NT>()
#endif
  double rr = 0.0;
  if (warp == kWSweepA || warp == kWSweepB) {
    // ------------------------------------------------------------------ sweeps gA = 2 warp + 1 (row t - L - 2 gA) and
    //                                                                    gB = gA + 1 (row t - L - 2 gA - 2)
    const int gA = (warp == kWSweepA) ? 1 : 3;
    const float* Sin = S + ((size_t)(gA - 1) * 2 * kSLanes + lane) * CP;
    float* Sout = S + ((size_t)(gA + 1) * 2 * kSLanes + lane) * CP;
    const float* Rl = R + (size_t)lane * CP;
    const float4* cl = coef + lane;
    int e0 = (1 - 2 * gA) & (kCoefSlots - 1);      // slot of entry t - 2 gA at t = 1 (same index for the R ring)
    // Register rings, indexed by compile-time step phases so that nothing is ever moved (the loop is unrolled over the
    // common period 6):
    //   set[q], q = t mod 3: coefficient set sweep A uses at step t -- [cy(C+1) | ninv(C)], cx of the row below, r --
    //     fetched BEFORE the previous step's barrier (entries <= t and the R slots <= t - 1 are complete by then), so only
    //     the previous stage's row is loaded on the critical path; sweep B at step t uses A's set of step t - 2 =
    //     set[(q+1) mod 3] (entries tau <= 0 are zero, so the ring starts as zeros);
    //   rA[q]: sweep A's result of step t; rA[(q+2) mod 3] = of step t - 1 (A's W operand, B's E operand),
    //     rA[(q+1) mod 3] = of step t - 2 (B's N operands);
    //   EA[p], rB[p], p = t mod 2: the previous stage's row loaded at step t (this step's E, next step's N operands) and
    //     sweep B's result (next step's W operand).
    float cyn[3][2 * C + 1], cxE[3][C], rv[3][C], rA[3][C], EA[2][C], rB[2][C], cxWB[C];
#pragma unroll
    for (int q = 0; q < 3; q++) {
#pragma unroll
      for (int c = 0; c < C; c++) { cxE[q][c] = 0.f; rv[q][c] = 0.f; rA[q][c] = 0.f; }
#pragma unroll
      for (int c = 0; c < 2 * C + 1; c++) cyn[q][c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < C; c++) { EA[0][c] = 0.f; EA[1][c] = 0.f; rB[0][c] = 0.f; rB[1][c] = 0.f; cxWB[c] = 0.f; }
    auto fetch = [&](auto q_c) {                    // the set of the step with phase q
      constexpr int q = decltype(q_c)::value;
      const int e1 = (e0 + 1) & (kCoefSlots - 1);
      ld_entry<F_CY, F_CX>(cl + e0 * ES, cyn[q]);
      ld_entry<F_CX, F_END>(cl + e1 * ES, cxE[q]);
      ld_block<C>(Rl + e0 * 32 * CP, rv[q]);
    };
    fetch(std::integral_constant<int, 1>{});        // t = 1
    auto step = [&](auto t_c) {
      constexpr int tt = decltype(t_c)::value;      // t mod 6
      constexpr int par = tt & 1, q = tt % 3, q1 = (q + 2) % 3, q2 = (q + 1) % 3;
      float SlA = __shfl_up_sync(0xffffffffu, rA[q1][C - 1], 1);
      float SlB = __shfl_up_sync(0xffffffffu, rB[par ^ 1][C - 1], 1);
      float NxB = __shfl_down_sync(0xffffffffu, rA[q1][0], 1);   // A's row t - L - 2 gA - 2, column 0 of lane L+1 (its last step's result)
      const float* Sp = Sin + (par ^ 1) * kSLanes * CP;
      ld_block<C>(Sp, EA[par]);
      const float NxA = Sp[CP];                     // column 0 of lane L+1 (lane 32 = zero pad)
      SlA = (lane == 0) ? 0.f : SlA;
      SlB = (lane == 0) ? 0.f : SlB;
      NxB = (lane == 31) ? 0.f : NxB;               // (the zero pad of the stage buffers)
      float resA[C];
#pragma unroll
      for (int c = 0; c < C; c++) {
        const float Sop = (c == 0) ? SlA : resA[c == 0 ? 0 : c - 1];
        const float Nop = (c == C - 1) ? NxA : EA[par ^ 1][c == C - 1 ? c : c + 1];
        resA[c] = (rA[q1][c] * cxE[q1][c] + EA[par][c] * cxE[q][c] + Sop * cyn[q][c] + Nop * cyn[q][c + 1] - rv[q][c]) * cyn[q][C + 1 + c];
      }
      float resB[C];
#pragma unroll
      for (int c = 0; c < C; c++) {                 // sweep B: E = A's result of the last step, N = A's result of two steps ago
        const float Sop = (c == 0) ? SlB : resB[c == 0 ? 0 : c - 1];
        const float Nop = (c == C - 1) ? NxB : rA[q2][c == C - 1 ? c : c + 1];
        resB[c] = (rB[par ^ 1][c] * cxWB[c] + rA[q1][c] * cxE[q2][c] + Sop * cyn[q2][c] + Nop * cyn[q2][c + 1] - rv[q2][c]) * cyn[q2][C + 1 + c];
      }
      st_block<C>(Sout + par * kSLanes * CP, resB);
#pragma unroll
      for (int c = 0; c < C; c++) { rA[q][c] = resA[c]; rB[par][c] = resB[c]; cxWB[c] = cxE[q2][c]; }
      e0 = (e0 + 1) & (kCoefSlots - 1);
      fetch(std::integral_constant<int, q2>{});     // (phase of step t + 1; its old content was sweep B's set of this step)
      RLFC_STEP_SYNC();
    };
    for (int t = 1; t <= t_end; t += 6) {
      step(std::integral_constant<int, 1>{});
      step(std::integral_constant<int, 2>{});
      step(std::integral_constant<int, 3>{});
      step(std::integral_constant<int, 4>{});
      step(std::integral_constant<int, 5>{});
      step(std::integral_constant<int, 0>{});
    }
  } else if (warp == kWResA || warp == kWResB) {
    // ------------------------------------------------------------------ residual increment, row i5 = t - L - 10
    // (mode 3: warp kWResA takes columns [0, CA) of every lane's block, warp kWResB columns [CA, C))
    if (XMODE == 3) {
      constexpr int CA = RLFC_ROWS_RES_SPLIT ? (C + 1) / 2 : C;
      auto run = [&](auto half_c) {
        constexpr int CLO = decltype(half_c)::value ? CA : 0, CHI = decltype(half_c)::value ? C : CA;
        float dC[C], dW[C], dEp[C], cxW[C];
#pragma unroll
        for (int c = 0; c < C; c++) { dC[c] = 0.f; dW[c] = 0.f; dEp[c] = 0.f; cxW[c] = 0.f; }
        int i5 = 1 - lane - kStageLag;
        const float* Sin = S + ((size_t)4 * 2 * kSLanes + lane) * CP;
        const float* Rl = R + (size_t)lane * CP;
        const float4* cl = coef + lane;
        int e0 = (1 - kStageLag) & (kCoefSlots - 1);
        float* rout = r + ((ptrdiff_t)(1 - kStageLag) * 32 + lane) * CP;   // skewed entry t - 10 of this lane
        // which of this lane's columns are the first / last column of the level
        const int cfirst = (lane == 0) ? 0 : -1;
        const int clast = (lane == nl - 1) ? (mj - 1) - C * lane : -1;
        auto step = [&](auto par_c) {
          constexpr int par = decltype(par_c)::value;
          const int e1 = (e0 + 1) & (kCoefSlots - 1);
          const float* Sp = Sin + (par ^ 1) * kSLanes * CP;
          float dE[C], cxE[C], cy[C + 1], rv[C], rN[C];
          ld_block<C>(Sp, dE);
          const float Nx = Sp[CP];
#pragma unroll
          for (int c = 0; c < C; c++) { dW[c] = dC[c]; dC[c] = dEp[c]; dEp[c] = dE[c]; }
          float Sl = 0.f;
          if (CLO == 0) Sl = __shfl_up_sync(0xffffffffu, dW[C - 1], 1);
          ld_entry<F_CY, F_NINV>(cl + e0 * ES, cy);
          ld_entry<F_CX, F_END>(cl + e1 * ES, cxE);
          ld_block<C>(Rl + e0 * 32 * CP, rv);
          const bool rowok = (unsigned)(i5 - 1) < (unsigned)ni;
          const bool top = i5 == 1, bot = i5 == ni;
#pragma unroll
          for (int c = CLO; c < CHI; c++) {
            const float d0 = dC[c];
            const float dg = -(cxW[c] + cxE[c] + cy[c] + cy[c + 1]);   // diagonal = -sumd, PoissonMatrix.pde:46-48
            const float w_ = top ? d0 : dW[c];
            const float e_ = bot ? d0 : dE[c];
            const float s_ = (c == cfirst) ? d0 : ((c == 0) ? Sl : dC[c == 0 ? 0 : c - 1]);
            const float n_ = (c == clast) ? d0 : ((c == C - 1) ? Nx : dC[c == C - 1 ? c : c + 1]);
            const float Ad = d0 * dg + w_ * cxW[c] + e_ * cxE[c] + s_ * cy[c] + n_ * cy[c + 1];   // PoissonMatrix.pde:56-61
            const bool ok = rowok && C * lane + c < mj;
            rN[c] = ok ? rv[c] - Ad : 0.f;
            const float prod = rN[c] * rN[c];            // float product, double accumulation (Field.pde:304-307)
            rr += (double)prod;
          }
#pragma unroll
          for (int c = 0; c < C; c++) cxW[c] = cxE[c];
          if (rowok) {                                   // this warp's share of the contiguous 32*CP-float block of the step
#pragma unroll
            for (int c = CLO; c < CHI; c++) {
              if (!(c & 1) && c + 1 < CHI) *reinterpret_cast<float2*>(rout + c) = make_float2(rN[c], rN[c + 1]);
              else if ((c & 1) && c > CLO) {}            // (second half of a pair)
              else rout[c] = rN[c];
            }
          }
          e0 = e1;
          i5++;
          rout += 32 * CP;
          RLFC_STEP_SYNC();
        };
        for (int t = 1; t <= t_end; t += 2) {
          step(std::integral_constant<int, 1>{});
          step(std::integral_constant<int, 0>{});
        }
      };
      if (warp == kWResA) run(std::integral_constant<int, 0>{});
      else if (CA < C) run(std::integral_constant<int, 1>{});
      else { for (int t = 1; t <= t_end; t++) RLFC_STEP_SYNC(); }
    } else {
      for (int t = 1; t <= t_end; t++) RLFC_STEP_SYNC();
    }
  } else if (warp == kWStage0) {
    // ------------------------------------------------------------------ stage 0 (row t - L) + loader
    int i0 = 1 - lane;
    int slot0 = ring_mod(i0);
    float* Sout = S + (size_t)lane * CP;
    float* Rl = R + (size_t)lane * CP;
    const float4* cl = coef + lane;
    for (int t = 1; t <= t_end; t++) {
      const int par = t & 1, e0 = t & (kCoefSlots - 1);
      float ninv[C], rv[C], d0[C];
      {
      ld_entry<F_NINV, F_CX>(cl + e0 * ES, ninv);
      if (SKEW) {
        ld_block<C>(Rl + e0 * 32 * CP, rv);          // entry t of the skewed residual (zero outside the domain)
      } else {
        const bool rowok = (unsigned)(i0 - 1) < (unsigned)ni;
        const float* rrow = rring + (size_t)slot0 * P + j0;
#pragma unroll
        for (int c = 0; c < C; c++) rv[c] = (rowok && C * lane + c < mj) ? rrow[c] : 0.f;
        st_block<C>(Rl + e0 * 32 * CP, rv);
      }
#pragma unroll
      for (int c = 0; c < C; c++) d0[c] = rv[c] * (-ninv[c]);   // MG.pde:80
      st_block<C>(Sout + par * kSLanes * CP, d0);
      }
      i0++;
      slot0 = ring_inc(slot0);
      RLFC_STEP_SYNC();
    }
  } else if (warp == kWXinc) {
    // ------------------------------------------------------------------ x increment: row t - L - 9 (sweep 4's last row)
    int i6 = 1 - lane - 9;
    int slot = ring_mod(i6);
    const float* Sin = S + ((size_t)4 * 2 * kSLanes + lane) * CP;
    float* gtop = gbuf;
    float* gbot = gbuf + mj;
    float* gleft = gbuf + 2 * mj;
    float* gright = gbuf + 2 * mj + ni;
    const int clastx = (lane == nl - 1) ? (mj - 1) - C * lane : -1;   // this lane's column that is the level's last one
    for (int t = 1; t <= t_end; t++) {
      const int par = t & 1;
      float d[C], xv[C];
      ld_block<C>(Sin + (par ^ 1) * kSLanes * CP, d);
      const bool rowok = (unsigned)(i6 - 1) < (unsigned)ni;
      float* xrow = xring + (size_t)slot * P + j0;
      if (XMODE != 1) {
#pragma unroll
        for (int c = 0; c < C; c++) xv[c] = (rowok && C * lane + c < mj) ? xrow[c] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < C; c++) {
        const float xn = (XMODE == 1) ? 0.f + d[c] : xv[c] + d[c];
        if (rowok && C * lane + c < mj) xrow[c] = xn;
      }
      if (XMODE == 3) {
        // boundary values of d feed the ghost cells of x after the sweep (x.plusEq(d) runs over all cells).  The first and
        // the last column are one predicated store each; the top / bottom rows are a branch a lane takes twice per call
        float dl = d[0];
#pragma unroll
        for (int c = 1; c < C; c++) dl = (c == clastx) ? d[c] : dl;
        if (rowok && lane == 0) gleft[i6 - 1] = d[0];
        if (rowok && clastx >= 0) gright[i6 - 1] = dl;
        if (i6 == 1 || i6 == ni) {
#pragma unroll
          for (int c = 0; c < C; c++)
            if (C * lane + c < mj) {
              if (i6 == 1) gtop[j0 + c - 1] = d[c];
              if (i6 == ni) gbot[j0 + c - 1] = d[c];
            }
        }
      }
      if (!BULK_WO) {  // write-out of row t - lag: every lane's x increment passed it at step t - 2
        const int w = t - lag;
        if (w >= 1 && w <= ni) {
          const float* xs = xring + (size_t)ring_mod(w) * P;
          float xo[C];
#pragma unroll
          for (int c = 0; c < C; c++) {
            const int j = 1 + lane + 32 * c;
            xo[c] = (j <= mj) ? xs[j] : 0.f;
          }
          float* xg = x + (size_t)w * P + 1 + lane;
#pragma unroll
          for (int c = 0; c < C; c++)
            if (1 + lane + 32 * c <= mj) xg[32 * c] = xo[c];
        }
      } else {
        // (the loader warp writes finished rows out as ONE bulk copy each; make this step's ring writes
        // visible to the async proxy before the step barrier)
        fence_proxy_async();
      }
      i6++;
      slot = ring_inc(slot);
      RLFC_STEP_SYNC();
    }
  } else if (warp == kWLoader) {
    // ------------------------------------------------------------------ loader: bulk copies kPF steps ahead
    // ... and (BULK_WO) the write-out of finished x rows: row w = t - lag is complete in the x ring after step t - 2
    // (every lane's increment has passed it), ghost columns and pitch padding still as loaded, so the whole row goes
    // back with one shared->global bulk copy.  Its ring slot is refilled with row w + kRowRing at step
    // w + kRowRing - kPF > w + lag: the copies of earlier steps must have READ their rows before this step's loads.
    // (everything lane 0 needs per step is kept as running pointers / offsets: the loader shares a scheduler with a sweep warp)
    {
      int q = 1 + kPF;                                                  // entry / row fetched at step t: q = t + kPF
      unsigned qoff = (unsigned)ring_mod(q) * row_bytes;               // byte offset of row q's slot in the row rings
      unsigned woff = (unsigned)ring_mod(1 - lag) * row_bytes;         // ... of row t - lag's slot
      const unsigned ring_bytes = (unsigned)kRowRing * row_bytes;
      const char* g_r = reinterpret_cast<const char*>(r) + (size_t)q * (SKEW ? rsk_bytes : row_bytes);
      char* g_x = reinterpret_cast<char*>(x) + (size_t)q * row_bytes;
      char* g_w = reinterpret_cast<char*>(x) + (ptrdiff_t)(1 - lag) * (ptrdiff_t)row_bytes;
      const char* g_tab = reinterpret_cast<const char*>(tab + (size_t)(q + kTabFront) * ES);
      char* const s_r = reinterpret_cast<char*>(rring);
      char* const s_x = reinterpret_cast<char*>(xring);
      char* const s_R = reinterpret_cast<char*>(R);
      char* const s_c = reinterpret_cast<char*>(coef);
      const unsigned tx_norow = ent_bytes + (SKEW ? rsk_bytes : 0u);
      const unsigned tx_row = tx_norow + ((XMODE != 1 ? 1u : 0u) + (SKEW ? 0u : 1u)) * row_bytes;
      for (int t = 1; t <= t_end; t++) {
        if (lane == 0) {
        if (BULK_WO) bulk_wait_read0();
        fence_proxy_async();
        {
          unsigned long long* bar = bars + ((q - 1) & (kCoefSlots - 1));
          const bool row = q <= ni;
          mbar_expect_tx(bar, row ? tx_row : tx_norow);
          if (row) {
            if (!SKEW) bulk_g2s(s_r + qoff, g_r, row_bytes, bar);
            if (XMODE != 1) bulk_g2s(s_x + qoff, g_x, row_bytes, bar);
          }
          if (SKEW) bulk_g2s(s_R + (size_t)(q & (kRSlots - 1)) * rsk_bytes, g_r, rsk_bytes, bar);
          bulk_g2s(s_c + (size_t)(q & (kCoefSlots - 1)) * ent_bytes, g_tab, ent_bytes, bar);
        }
        if (BULK_WO) {
          const int w = t - lag;
          if (w >= 1 && w <= ni) {
            bulk_s2g(g_w, s_x + woff, row_bytes);
            bulk_commit();
          }
        }
        }
        q++;
        qoff = (qoff + row_bytes == ring_bytes) ? 0u : qoff + row_bytes;
        woff = (woff + row_bytes == ring_bytes) ? 0u : woff + row_bytes;
        g_r += SKEW ? rsk_bytes : row_bytes;
        g_x += row_bytes;
        g_w += row_bytes;
        g_tab += ent_bytes;
        mbar_wait(bars + (t & (kCoefSlots - 1)), ((unsigned)t >> kCoefShift) & 1u);   // entry / row t + 1 has landed
        RLFC_STEP_SYNC();
      }
    }
    // drain: every issued copy must have landed before the shared memory is reused
    for (int q = t_end + 2; q <= t_end + kPF; q++) mbar_wait(bars + ((q - 1) & (kCoefSlots - 1)), ((unsigned)(q - 1) >> kCoefShift) & 1u);
    if (BULK_WO && lane == 0) {   // ... and every written-out row must be in global memory before anybody reads x again
      bulk_wait0();
      fence_proxy_async_all();
    }
  } else {
    for (int t = 1; t <= t_end; t++) RLFC_STEP_SYNC();
  }
#undef RLFC_STEP_SYNC
#ifdef RLFC_ROWS_TIMING
  if (XMODE == 3 && blockIdx.x == 0 && lane == 0)
    printf("rows level0 C=%d warp %d: busy %lld cycles over %d steps = %lld per step\n", C, warp, busy_, t_end, busy_ / t_end);
#endif
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < kCoefSlots; k++) mbar_inval(bars + k);
  __syncthreads();
  return rr;
}

}  // namespace rlfc
