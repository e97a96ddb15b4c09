"""CPU models (numpy, float32, one rounding per operation) of the two smoother schedules introduced in round 2, checked against
a plain lexicographic Gauss-Seidel with the reference's statement order (MG.smooth, MG.pde:79-92):

* the row pipeline's sweep warps run TWO consecutive sweeps each (csrc/smooth_rows.cuh): the second sweep of a pair takes its
  coefficients from a three-deep register ring (the first sweep's sets of two steps ago) and its operands from the first
  sweep's last two results;
* the single-warp smoother of the small levels (csrc/smooth_tiny.cuh): lane L owns column L+1 and runs stage g on row
  t - L - 2g for all five stages in one step, every operand being a register of the same lane or of a neighbour lane from the
  previous step.

The models follow the kernels' index algebra (phases t mod 2 / t mod 3, lane skew, zero coefficients outside the domain), so
an edit that breaks a schedule shows up here without a GPU; the arithmetic itself is checked on the device against the oracle."""
import numpy as np
import pytest

F = np.float32


def make_level(ni, mj, seed):
    rng = np.random.default_rng(seed)
    n, m = ni + 2, mj + 2
    lx = rng.uniform(0.1, 0.3, size=(n + 1, m + 1)).astype(F)
    ly = rng.uniform(0.1, 0.3, size=(n + 1, m + 1)).astype(F)
    # coarse-level convention: boundary faces are 0 (MG.pde:120), so ghost values of d never contribute
    lx[1, :] = 0; lx[n - 1, :] = 0; ly[:, 1] = 0; ly[:, m - 1] = 0
    inv = np.ones((n, m), F)
    for i in range(1, n - 1):
        for j in range(1, m - 1):
            s = F(F(F(lx[i, j] + lx[i + 1, j]) + ly[i, j]) + ly[i, j + 1])
            inv[i, j] = F(-1.0) / s
    r = rng.normal(size=(n, m)).astype(F)
    r[0, :] = r[-1, :] = 0; r[:, 0] = r[:, -1] = 0
    return r, lx, ly, inv


def update(W, cxW, E, cxE, S, cy, N, cyN, rv, ninv):
    """d = (dW*lxW + dE*lxE + dS*lyS + dN*lyN - r) * (-inv), left to right, one float32 rounding per operation."""
    acc = F(W * cxW)
    acc = F(acc + F(E * cxE))
    acc = F(acc + F(S * cy))
    acc = F(acc + F(N * cyN))
    acc = F(acc - rv)
    return F(acc * ninv)


def reference_gs(r, lx, ly, inv, sweeps=4):
    n, m = r.shape
    d = (r * inv).astype(F)
    d[0, :] = d[-1, :] = 0; d[:, 0] = d[:, -1] = 0
    for _ in range(sweeps):
        for i in range(1, n - 1):
            for j in range(1, m - 1):
                d[i, j] = update(d[i - 1, j], lx[i, j], d[i + 1, j], lx[i + 1, j], d[i, j - 1], ly[i, j], d[i, j + 1], ly[i, j + 1],
                                 r[i, j], F(-inv[i, j]))
    return d


def coef_set(r, lx, ly, inv, i, j):
    """(lxW, lxE, lyS, lyN, -inv, r) of cell (i, j); zeros outside the interior (the kernels' zero table entries)."""
    n, m = r.shape
    if 1 <= i <= n - 2 and 1 <= j <= m - 2:
        return lx[i, j], lx[i + 1, j], ly[i, j], ly[i, j + 1], F(-inv[i, j]), r[i, j]
    return (F(0),) * 6


def tiny_model(r, lx, ly, inv):
    """smooth_tiny.cuh: one warp, lane L = column L+1, five stages per step."""
    n, m = r.shape
    ni, mj = n - 2, m - 2
    assert mj <= 32
    p = np.zeros((5, 32), F)                       # previous-step results of stages 0..4, per lane
    d = np.zeros_like(r)
    for t in range(1, ni + mj + 9 + 1):
        q = np.zeros((5, 32), F)
        for L in range(32):
            j, i0 = L + 1, t - L
            _, _, _, _, ninv0, r0 = coef_set(r, lx, ly, inv, i0, j)
            q[0, L] = F(r0 * F(-ninv0))            # stage 0: d = r * inv
            for g in range(1, 5):
                i = i0 - 2 * g
                lw, le, ls, ln, ninv, rv = coef_set(r, lx, ly, inv, i, j)
                S = p[g, L - 1] if L > 0 else F(0)
                N = p[g - 1, L + 1] if L < 31 else F(0)
                q[g, L] = update(p[g, L], lw, p[g - 1, L], le, S, ls, N, ln, rv, ninv)
            i4 = i0 - 8
            if 1 <= i4 <= ni and L < mj:
                d[i4, j] = q[4, L]
        p = q
    return d


def pair_model(r, lx, ly, inv, C):
    """smooth_rows.cuh: lane L owns columns C*L+1 .. C*L+C; warp k runs sweeps gA = 2k+1 (row t - L - 2gA) and gA+1 (two rows
    behind); register rings indexed by t mod 3 (coefficient sets, sweep A's results) and t mod 2 (previous stage's rows, sweep
    B's results); stage buffers S[stage][t mod 2] hand rows from stage 0 to warp 0 and from warp 0 to warp 1."""
    n, m = r.shape
    ni, mj = n - 2, m - 2
    nl = (mj + C - 1) // C
    assert nl <= 32
    NL = 32

    def sets(tau, L):                              # the entry of step tau for lane L: its row is tau - L
        return [coef_set(r, lx, ly, inv, tau - L, C * L + 1 + c) for c in range(C)]

    S = np.zeros((5, 2, NL + 1, C), F)             # [stage][parity][lane (+ zero 33rd)][column]
    ring = [[[[(F(0),) * 6 for _ in range(C)] for _ in range(NL)] for _ in range(3)] for _ in range(2)]   # [warp][q][lane][c]
    rA = np.zeros((2, 3, NL, C), F)
    EA = np.zeros((2, 2, NL, C), F)
    rB = np.zeros((2, 2, NL, C), F)
    cxWB = np.zeros((2, NL, C), F)
    for k in range(2):
        for L in range(NL):
            ring[k][1][L] = sets(1 - 2 * (2 * k + 1), L)           # fetch of t = 1
    d = np.zeros_like(r)
    t_end = (ni + nl + 10 + 5) // 6 * 6
    for t in range(1, t_end + 1):
        par, q, q1, q2 = t & 1, t % 3, (t + 2) % 3, (t + 1) % 3
        Snew = S.copy()
        for L in range(NL):                        # stage 0 (its own warp): row t - L
            for c, (_, _, _, _, ninv, rv) in enumerate(sets(t, L)):
                Snew[0, par, L, c] = F(rv * F(-ninv))
        for k in range(2):
            gA = 2 * k + 1
            newA = np.zeros((NL, C), F); newB = np.zeros((NL, C), F)
            for L in range(NL):
                EA[k, par, L] = S[gA - 1, par ^ 1, L]
                NxA = S[gA - 1, par ^ 1, L + 1, 0]
                SlA = rA[k, q1, L - 1, C - 1] if L > 0 else F(0)
                SlB = rB[k, par ^ 1, L - 1, C - 1] if L > 0 else F(0)
                NxB = rA[k, q1, L + 1, 0] if L < 31 else F(0)
                for c in range(C):                 # sweep A
                    lw, le, ls, ln, ninv, rv = ring[k][q][L][c]
                    cxW = ring[k][q1][L][c][1]     # lxE of the previous step's set = lxW of this row
                    Sop = SlA if c == 0 else newA[L, c - 1]
                    Nop = NxA if c == C - 1 else EA[k, par ^ 1, L, c + 1]
                    newA[L, c] = update(rA[k, q1, L, c], cxW, EA[k, par, L, c], le, Sop, ls, Nop, ln, rv, ninv)
                for c in range(C):                 # sweep B: the set of two steps ago, operands from sweep A's last two results
                    lw, le, ls, ln, ninv, rv = ring[k][q2][L][c]
                    Sop = SlB if c == 0 else newB[L, c - 1]
                    Nop = NxB if c == C - 1 else rA[k, q2, L, c + 1]
                    newB[L, c] = update(rB[k, par ^ 1, L, c], cxWB[k, L, c], rA[k, q1, L, c], le, Sop, ls, Nop, ln, rv, ninv)
            for L in range(NL):
                Snew[gA + 1, par, L] = newB[L]
                for c in range(C):
                    cxWB[k, L, c] = ring[k][q2][L][c][1]
                ring[k][q2][L] = sets(t + 1 - 2 * gA, L)            # fetch for step t + 1 (phase q2)
            rA[k, q] = newA
            rB[k, par] = newB
            if k == 1:                             # sweep 4's row of this step
                for L in range(NL):
                    i = t - L - 8
                    for c in range(C):
                        j = C * L + 1 + c
                        if 1 <= i <= ni and j <= mj:
                            d[i, j] = newB[L, c]
        S = Snew
    return d


@pytest.mark.parametrize("ni,mj", [(12, 6), (10, 24), (7, 32), (20, 3)])
def test_single_warp_schedule(ni, mj):
    r, lx, ly, inv = make_level(ni, mj, seed=ni * 100 + mj)
    ref = reference_gs(r, lx, ly, inv)
    got = tiny_model(r, lx, ly, inv)
    assert np.array_equal(got[1:-1, 1:-1], ref[1:-1, 1:-1])


@pytest.mark.parametrize("ni,mj,C", [(12, 12, 1), (14, 20, 2), (10, 33, 3), (9, 70, 3), (8, 40, 6)])
def test_two_sweeps_per_warp_schedule(ni, mj, C):
    r, lx, ly, inv = make_level(ni, mj, seed=ni * 1000 + mj * 10 + C)
    ref = reference_gs(r, lx, ly, inv)
    got = pair_model(r, lx, ly, inv, C)
    assert np.array_equal(got[1:-1, 1:-1], ref[1:-1, 1:-1])
