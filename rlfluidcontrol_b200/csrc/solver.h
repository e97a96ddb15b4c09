// solver.h -- device data layout and kernel launch interface of the batched BDIM solver.
//
// Layout in HBM (all fp32):
//   * every per-environment field is an (n x P) pitched array (i = x outer, j = y contiguous, ghost
//     ring included, P = m rounded up to 8 floats so rows are 32 B-sector aligned), and the batch is
//     struct-of-arrays: field[e] = base + e * stride, stride = n*P rounded up to 32 floats;
//   * static geometry / coefficient fields are stored once (shared by the whole batch, L2-resident).
// Velocity lives in three rotating buffers A (step-start u, doubles as u0), B (predictor result,
// doubles as `us`), C (corrector result); see BDIM.pde:79-107 for the data flow reproduced.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace rlfc {

constexpr int kMaxLevels = 16;

// Pre-skewed static coefficient tables of one level for the strip smoother (smooth_strip.cuh):
// entry [(k*Tsk + tau + kSkewPad)*32 + lane] belongs to cell (i = tau - lane + 1, j = 32k + lane + 1).
struct SkewLevel {
  const float4* A;          // (lx[i][j], lx[i+1][j], ly[i][j], ly[i][j+1])
  const float2* nd;         // (-inv[i][j], diag[i][j])
  int nstrips, Tsk;
};

// Pre-skewed static coefficient table of one level for the row-pipelined smoother (smooth_rows.cuh):
// entry tau (stored at index tau + kTabFront) holds, for lane l, the coefficients of row tau - l of the
// lane's C columns j0 = C*l+1 .. C*l+C as the float stream [ly[row][j0+c] (C+1) | -inv[row][j0+c] (C) |
// lx[row][j0+c] (C)], packed into K = ceil((3C+1)/4) float4 vectors that are
// lane-contiguous: vector k of lane l of entry e sits at T[(e*K + k)*32 + l].
struct RowTab {
  const float4* T;
  int C, K, entries, copies;   // `copies` identical tables back to back (L2-slice load spreading)
};

#define IDX(i, j) ((i) * P + (j))

// Chained strip smoother (smooth_chain.cuh): per level, the columns are cut into strips of 32 (lane L of strip s owns
// column j = 32 s + L + 1) and every array the sweeps touch is stored STRIP-SKEWED: element (s, tau, L) at
// (s*T + tau)*32 + L holds the cell (i = tau - L, j), so that the cells a warp works on in one step form one
// contiguous 128-byte line.  Entries 0 .. T-1 with T = ni + 34; elements outside the interior are zero.
struct ChainLevel {
  int on;                   // 1 = this level is smoothed by the chained strip smoother
  int NS, T;                // strips, entries per strip
  int wpb, nb;              // strips (warps) per CTA, CTAs per sweep
  size_t sk_stride;         // elements per environment of a skewed array = NS*T*32
  const float4* ct;         // static coefficients {lx[i+1][j], ly[i][j], ly[i][j+1], -inv[i][j]}   [NS][T][32]
  float* rsk;               // [B] residual entering smooth(4)                                     (4 B / element)
  uint2* dsk[5];            // [B] iterate after sweep g = 0..4 as {value bits, launch tag}         (8 B / element)
  unsigned long long* ticket;   // [1] CTA tickets of this level's sweep launches (role order = start order)
  unsigned tag_hi;          // high bits of the launch tag (unique per ticket counter)
  // smooth_chain3.cuh: compact copies of every strip's lane-31 (S edge) and lane-0 (N edge) results, indexed by the
  // CONSUMER's step (+ kC2EdgePad), array (2 g + which) of sweep g at edge + (2 g + which) * edge_arr, [B][NS][TE]
  uint2* edge;
  size_t edge_arr;
  int TE;
  int s0, ns_loc;           // strips s0 .. s0 + ns_loc - 1 are swept by this device (slab mode; otherwise 0, NS)
};

struct DevLevel {
  SkewLevel sk;
  RowTab rt;
  int n, m, P;              // dims incl. ghosts, pitch
  size_t stride;            // per-env stride (floats) of r/x/d at this level
  const float *lx, *ly, *inv, *diag;   // static, pitched
  float *r, *r2, *x, *d;    // per-env batch arrays (level 0: x = p; r2 = ping-pong residual)
  float *w;                 // level 0, wavefront smoother only: scratch for the Gauss-Seidel iterate
  int wave;                 // 1 = this level is smoothed by the wavefront fallback (smooth_wave.cuh)
  ChainLevel ch;            // chained strip smoother (wide levels; smooth_chain.cuh)
};

struct BandFace {           // a face where the BDIM blend differs from the identity
  int i, j;
  float del, del1, wnx, wny;
};

struct SamplePt { int i, j; float s, t; };
struct ForcePt { int i, j; float s, t, l, nx, ny; };

// Per-env scalar state (struct of arrays on the device)
struct EnvScalars {
  float *xi;        // [B][2] current action (xi1, xi2)
  float *t;         // [B]    AFCCylinder.t
  float *flow_t;    // [B]    BDIM.t: grid-unit time, += dt per solver step (BDIM.pde:106), set from the checkpoint
  float *force;     // [B][2] last raw force (-pressForce)
  float *probes;    // [B][32]
  int   *callLearn; // [B]    draw() accumulators (clientCFD.pde:11-13)
  float *Cd, *Cl;   // [B]
  float *obs;       // [B][2] last produced (Cl, Cd)
  int   *frozen;    // [B]    env takes no part in the solver steps of this call: it has emitted its observation and waits
                    //        for the next action (clientCFD.pde:44-54), or its episode is over (clientCFD.pde:36)
  int   *non_finite;// [B]    sticky: a non-finite force was produced (diverged environment)
  int   *n_running; // [1]    environments not frozen, counted by k_emit_obs
  int   *active;    // [B]    MG: env still iterating
  int   *iters;     // [B][2] MG iterations of the last predictor/corrector solve
  double *rr_part;  // [B][rr_blocks] partial sums of r.r
  float *psum;      // [B]    serial interior sum of p
  int   *any_active;// [1]
};

// Slab mode (BASELINE config 5): one domain advanced by several devices that share every array through one virtual
// address range (vmm.h).  A device owns a contiguous range of a kernel's row blocks (plain arrays) or strips (skewed
// arrays); kernels are launched with their full grid on every device and blocks outside the device's range exit.
struct SlabBarrier {
  unsigned* count;          // [1] arrivals of all devices, all barriers (memory of device 0, mapped everywhere)
  unsigned* epoch;          // [1] this device's barrier count (its own memory)
  unsigned n;               // devices
};

struct SolverParams {
  int slab_n, slab_rank;    // devices sharing the domain (0/1 = not in slab mode), this device's index
  int B;                    // environments in the batch
  int n, m, P;              // level-0 dims
  size_t stride;            // level-0 per-env stride
  float dt, nu, dRD;        // dRD = dR*D (AFCCylinder.pde:48)
  float dt_over_res;        // dt/resolution (AFCCylinder.pde:56)
  float action_scale;       // 5
  float inv_cells;          // (float)((n-2)*(m-2))
  float mg_tol;
  int   nlevels;
  int   coarse_strips;      // max strip count over levels >= 1 (warps of k_mg_coarse)
  int   chain_v;            // sweep kernel generation: 1 = smooth_chain.cuh, 3 (default) = smooth_chain3.cuh
  int   chain_levels;       // levels 0 .. chain_levels-1 run as grid-wide kernels with the chained strip smoother; the
                            // one-CTA-per-env coarse kernel starts at level max(chain_levels, 1)
  double *rr_chain;         // [B][rr_chain_n] per-CTA partial sums of r.r of the level-0 chain increment
  unsigned *rr_count;       // [B] CTAs of the increment kernel that have delivered their partial sum
  int   rr_chain_n;
  int   tiny;               // 1 = levels of at most 32 columns are smoothed by one warp in registers (smooth_tiny.cuh; RLFC_TINY=0: row pipeline)
  int   resid_march;        // 1 = k_resid_down0_march (default), 0 = the shared-memory tile version (RLFC_RESID=tile)
  int   fast_bc;            // 1 = two-phase setBC kernels (no band face on the lines setBC reads; grid fits one CTA)
  int   use_rows;           // 1 = row-pipelined smoother (smooth_rows.cuh), 0 = strip smoother (smooth_strip.cuh)
  int   resolution, substeps, mg_max_iters;
  float init_time, episode_time;
  DevLevel lev[kMaxLevels];
  // static level-0 fields
  const float *c_x, *c_y;
  const float *w1_x, *w2_x, *ry1_x, *ry2_x, *w1_y, *w2_y, *rx1_y, *rx2_y;
  const BandFace *band_x, *band_y;
  int nband_x, nband_y;
  float *band_tmp;          // [B][nband_x + nband_y]
  float *rsk;               // [B][rsk_stride] level-0 residual in the smoother's skewed layout (smooth_rows.cuh)
  size_t rsk_stride;
  const ForcePt *force_pts; int nforce;
  const SamplePt *probe_pts; int nprobe;
  int rr_blocks;            // number of per-env partial sums written by the increment kernel
  // Field.sum as segment summaries (exact_sum.cuh); xs_recs == nullptr selects the plain serial chain (k_psum)
  double *xs_ctot;          // [B][xs_nchunks] chunk totals, valid when xs_cflag == xs_epoch + 1
  unsigned *xs_cflag;       // [B][xs_nchunks]
  unsigned *xs_rflag;       // [B][xs_nchunks] == xs_epoch + 1 once the chunk's batch records are written
  unsigned *xs_epoch;       // [B] passes completed
  unsigned *xs_recs;        // [B][xs_nbatches][192] batch records (32 summaries condensed)
  unsigned *xs_blk;         // [B][ceil(xs_nbatches/32)][8][32] blocks of 32 records condensed (large domains; else nullptr)
  double *xs_pred, *xs_inc, *xs_corr;   // [B][xs_nbatches] large domains: first-pass prediction at every batch start, the batches'
                            // float increments, and the correction the second table pass adds (k_xsum_refine)
  int xs_nseg, xs_nchunks, xs_nbatches;
  int xs_passes;            // table passes of a large domain's Field.sum: 2, or 3 from 8 Mi cells on (RLFC_XS_PASSES)
  int xs_flags;             // cross-check switches: bit 0 = redo batches as plain additions (RLFC_XS_REDO=serial)
  int *xs_stats;            // [B][8] counters of the last serial pass (rlfc_env_field_sum_stats)
  EnvScalars sc;
};

#ifdef __CUDACC__
// true = block `blk` of `nblk` (a kernel's row-block or strip index) belongs to another device
__device__ __forceinline__ bool slab_skip(unsigned blk, unsigned nblk, int rank, int n) {
  if (n <= 1) return false;
  const unsigned lo = (unsigned)((unsigned long long)nblk * (unsigned)rank / (unsigned)n);
  const unsigned hi = (unsigned)((unsigned long long)nblk * (unsigned)(rank + 1) / (unsigned)n);
  return blk < lo || blk >= hi;
}
#endif

// every device of a slab-mode handle waits here for all the others (between dependent kernels)
int launch_slab_barrier(const SlabBarrier& b, cudaStream_t st);

// opt-in shared-memory attributes of the strip kernels; call once per handle before the first launch / capture
int configure_kernels(const SolverParams& P);
// CTAs of the level-0 chain increment kernel per environment (= per-env partial sums of r.r it writes)
int chain_incr_blocks(int ni, int NS);

// ---- launch wrappers (solver_kernels.cu).  All enqueue on `st`; return number of launches. ----
int launch_advdif(const SolverParams& P, const float* srcx, const float* srcy, const float* u0x, const float* u0y,
                  float* dstx, float* dsty, cudaStream_t st);
int launch_band_bc(const SolverParams& P, float* ux, float* uy, cudaStream_t st);
int launch_residual(const SolverParams& P, const float* ux, const float* uy, float* r, int which, cudaStream_t st);
// first MG iteration fused: residual + level-0 smooth(0)/increment/restriction; p_in -> p_out (different buffers)
int launch_resid_down0(const SolverParams& P, const float* ux, const float* uy, const float* p_in, float* p_out, float* r_out,
                       int which, cudaStream_t st);
// projection tail fused: p_out = p_in + shift (different buffers), u -= c * grad(p_in + shift)
int launch_project_shift(const SolverParams& P, const float* p_in, float* p_out, float* ux, float* uy, cudaStream_t st);
// corrector (P.fast_bc only): projection + Heun average u_out = (u + us)/2 away from the boundary lines, then u.setBC and
// the Heun average on the remaining zone
int launch_project_shift_heun(const SolverParams& P, const float* p_in, float* p_out, float* ux, float* uy, const float* usx,
                              const float* usy, float* uox, float* uoy, cudaStream_t st);
int launch_bc_heun(const SolverParams& P, float* ux, float* uy, const float* usx, const float* usy, float* uox, float* uoy,
                   cudaStream_t st);
// one MG iteration (V-cycle + smooth(4)) on active envs = down0, coarse, up0, smooth0;
// r_in/r_out are the level-0 ping-pong residual buffers
int launch_mg_down0(const SolverParams& P, const float* r_in, float* r_out, cudaStream_t st);
int launch_mg_coarse(const SolverParams& P, cudaStream_t st);
// the pieces launch_mg_coarse / launch_mg_up0 / launch_smooth0 are made of when levels use the chained strip smoother
int launch_chain_down(const SolverParams& P, int level, cudaStream_t st);
int launch_coarse_cta(const SolverParams& P, cudaStream_t st);
int launch_chain_up(const SolverParams& P, int level, const float* r, cudaStream_t st);
int launch_chain_sweeps(const SolverParams& P, int level, cudaStream_t st);
int launch_chain_incr(const SolverParams& P, int level, float* r_out, int which, cudaStream_t st);
int launch_mg_up0(const SolverParams& P, float* r, cudaStream_t st);
// rows smoother only: plain level-0 residual <- skewed residual (needed before a further MG iteration's down0)
int launch_unskew_r(const SolverParams& P, float* r, cudaStream_t st);
int launch_smooth0(const SolverParams& P, const float* r_in, float* r_out, int which, cudaStream_t st);
// last node of the MG iteration body inside a CUDA-graph WHILE node: cond = any env still active
int launch_loopcond(const SolverParams& P, unsigned long long cond_handle, cudaStream_t st);
int launch_psum(const SolverParams& P, cudaStream_t st);
int launch_psum_tables(const SolverParams& P, cudaStream_t st);   // (the two kernels of launch_psum, segment-summary mode)
int launch_psum_pass(const SolverParams& P, cudaStream_t st);
// same, with the serial pass on `side` (forked from / joined to `st` through the two events) so that it runs WHILE the
// table kernel produces the records it consumes chunk by chunk; for stream capture (parallel branches of the graph)
int launch_psum_overlapped(const SolverParams& P, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join);
int launch_project_u(const SolverParams& P, float* ux, float* uy, cudaStream_t st);
int launch_shift_p(const SolverParams& P, cudaStream_t st);
int launch_bc(const SolverParams& P, float* ux, float* uy, cudaStream_t st);
int launch_heun(const SolverParams& P, const float* ucx, const float* ucy, const float* ubx, const float* uby,
                float* uax, float* uay, cudaStream_t st);
// force + probes + time advance; accumulate != 0 applies the clientCFD.draw() accumulation
int launch_force(const SolverParams& P, int accumulate, cudaStream_t st);
// BDIM.checkCFL of every environment's current velocity: d_dt[B] = min(1/(max(|ux|+|uy|) + 3 nu), 1)
int launch_check_cfl(const SolverParams& P, const float* ux, const float* uy, float* d_dt, cudaStream_t st);
// mode 0 (single solver steps): xi = actions (NULL keeps xi), every env runs.  mode 1 (RL step): an env takes its
// action only where the reference would have asked for one (t > init_time and at a callLearn boundary); envs whose
// episode is over (t >= episode_time) are frozen for the call
int launch_set_actions(const SolverParams& P, const float* d_actions, int mode, cudaStream_t st);
int launch_emit_obs(const SolverParams& P, const float* d_actions, float* d_obs, float* d_reward, int* d_done,
                    cudaStream_t st);

}  // namespace rlfc
