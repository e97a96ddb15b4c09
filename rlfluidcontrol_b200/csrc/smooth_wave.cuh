// smooth_wave.cuh -- exact lexicographic Gauss-Seidel smoothing (MG.smooth, MG.pde:79-97) by anti-diagonal
// wavefronts in global memory: the size-independent fallback for levels too wide for the row pipeline
// (smooth_rows.cuh: at most 8 columns per lane on level 0, 4 on coarse levels).  Slow (one barrier per
// diagonal), but it runs a 2048 x 1024 domain, which the pipeline cannot.
//
// Cells with equal i + j are independent in a lexicographic sweep: (i,j) needs NEW values at (i-1,j), (i,j-1)
// (diagonal k-1) and OLD values at (i+1,j), (i,j+1) (diagonal k+1).  The four sweeps run in place and
// pipelined: in phase T sweep s updates diagonal T - 2s, which only reads diagonals at odd distance, all of
// them finished in phase T - 1 by the sweeps that own them.
#pragma once
#include <cuda_runtime.h>

#include "solver.h"

namespace rlfc {

// XMODE: 1  x = 0 + d        (coarsest level)          2  x = x + d
//        3  level 0: d.setBC, x += d on all cells, r_out = r - A d, returns this thread's share of r.r
// Called by all threads of the CTA.  `d` is a scratch array of the level's size (distinct from r, x, r_out).
template <int XMODE>
__device__ __forceinline__ double wave_smooth(const DevLevel& L, const float* __restrict__ r, float* __restrict__ x,
                                              float* __restrict__ d, float* __restrict__ r_out, int its) {
  const int n = L.n, m = L.m, P = L.P, ni = n - 2, mj = m - 2;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* __restrict__ lx = L.lx;
  const float* __restrict__ ly = L.ly;
  const float* __restrict__ inv = L.inv;
  // d = r * inv (MG.pde:80); ghost cells of d act as 0 during the sweeps (their products with the boundary
  // coefficients are +-0 on every level)
  for (int k = tid; k < n * P; k += nt) {
    const int i = k / P, j = k - i * P;
    d[k] = (i >= 1 && i <= ni && j >= 1 && j <= mj) ? r[k] * inv[k] : 0.f;
  }
  __syncthreads();
  const int kd_max = ni + mj;
  for (int T = 2; T <= kd_max + 2 * (its - 1); T++) {
    // flatten (sweep, cell on its diagonal) over the threads
    int cnt[8], lo[8], tot = 0;
#pragma unroll 1
    for (int s = 0; s < its; s++) {
      const int kd = T - 2 * s;
      lo[s] = max(1, kd - mj);
      const int hi = min(ni, kd - 1);
      cnt[s] = (kd >= 2 && kd <= kd_max && hi >= lo[s]) ? hi - lo[s] + 1 : 0;
      tot += cnt[s];
    }
    for (int w = tid; w < tot; w += nt) {
      int s = 0, o = w;
      while (o >= cnt[s]) { o -= cnt[s]; s++; }
      const int i = lo[s] + o, j = (T - 2 * s) - i, k = IDX(i, j);
      // d = -(dW*lxW + dE*lxE + dS*lyS + dN*lyN - r) * inv, MG.pde:85-86 ((-a)*b == a*(-b) bitwise)
      d[k] = (d[k - P] * lx[k] + d[k + P] * lx[k + P] + d[k - 1] * ly[k] + d[k + 1] * ly[k + 1] - r[k]) * (-inv[k]);
    }
    __syncthreads();
  }
  double rr = 0.0;
  if (XMODE == 3) {
    // d.setBC (ghost = adjacent interior value), x += d on all cells, r -= A d on the interior (MG.pde:90-97)
    for (int k = tid; k < n * m; k += nt) {
      const int i = k / m, j = k - i * m;
      const int ci = min(max(i, 1), ni), cj = min(max(j, 1), mj);
      x[IDX(i, j)] += d[IDX(ci, cj)];
      if (i >= 1 && i <= ni && j >= 1 && j <= mj) {
        const int q = IDX(i, j);
        const float dc = d[q];
        const float dW = d[IDX(max(i - 1, 1), j)], dE = d[IDX(min(i + 1, ni), j)];
        const float dS = d[IDX(i, max(j - 1, 1))], dN = d[IDX(i, min(j + 1, mj))];
        const float Ad = dc * L.diag[q] + dW * lx[q] + dE * lx[q + P] + dS * ly[q] + dN * ly[q + 1];   // PoissonMatrix.pde:56-61
        const float rN = r[q] - Ad;
        r_out[q] = rN;
        const float prod = rN * rN;                    // float product, double accumulation (Field.pde:304-307)
        rr += (double)prod;
      }
    }
  } else {
    for (int k = tid; k < ni * mj; k += nt) {
      const int i = 1 + k / mj, j = 1 + (k - (k / mj) * mj), q = IDX(i, j);
      x[q] = (XMODE == 1) ? 0.f + d[q] : x[q] + d[q];
    }
  }
  __syncthreads();
  return rr;
}

}  // namespace rlfc
