"""Index/timing model of the row-pipelined Gauss-Seidel smoother (csrc/smooth_rows.cuh), in numpy.

Simulates the exact dataflow of the kernel -- lanes owning C columns, one warp per sweep, the two-slot stage
buffers, carried registers, write-out lag -- and checks it against a plain lexicographic Gauss-Seidel
(MG.smooth semantics) plus the level-0 increment stage.  Same operation order => identical results.
Design aid only; nothing imports it."""
import numpy as np

LAG = 10


def reference(r, lx, ly, inv, diag, x, sweeps=4):
    n, m = r.shape
    d = r * inv
    d[0, :] = d[-1, :] = 0; d[:, 0] = d[:, -1] = 0     # ghost operands act as 0 (their coefficients are 0 / r ghost is 0)
    for _ in range(sweeps):
        for i in range(1, n - 1):
            for j in range(1, m - 1):
                d[i, j] = -(d[i - 1, j] * lx[i, j] + d[i + 1, j] * lx[i + 1, j] + d[i, j - 1] * ly[i, j] + d[i, j + 1] * ly[i, j + 1] - r[i, j]) * inv[i, j]
    # d.setBC (btype 0)
    d[0, :] = d[1, :]; d[-1, :] = d[-2, :]
    d[:, 0] = d[:, 1]; d[:, -1] = d[:, -2]
    xn = x + d
    rn = r.copy()
    for i in range(1, n - 1):
        for j in range(1, m - 1):
            rn[i, j] = r[i, j] - (d[i, j] * diag[i, j] + d[i - 1, j] * lx[i, j] + d[i + 1, j] * lx[i + 1, j] + d[i, j - 1] * ly[i, j] + d[i, j + 1] * ly[i, j + 1])
    return d, xn, rn


def pipelined(r, lx, ly, inv, diag, x, C):
    n, m = r.shape
    ni, mj = n - 2, m - 2
    nl = (mj + C - 1) // C
    NL = 32
    lag = nl + LAG
    t_end = ni + lag

    def cell(i, j):
        return 1 <= i <= ni and 1 <= j <= mj

    # table(tau, lane): cx[C], cy[C+1], ninv[C], diag[C] of row tau - lane
    def entry(tau, L):
        row = tau - L
        j0 = L * C + 1
        cx = np.zeros(C); cy = np.zeros(C + 1); ninv = np.zeros(C); dg = np.zeros(C)
        for c in range(C):
            j = j0 + c
            if 1 <= row <= ni + 1 and j <= mj:
                cx[c] = lx[row, j]
            if cell(row, j):
                ninv[c] = -inv[row, j]; dg[c] = diag[row, j]
        for c in range(C + 1):
            j = j0 + c
            if 1 <= row <= ni and j <= mj + 1:
                cy[c] = ly[row, j]
        return cx, cy, ninv, dg

    def rval(i, j):
        return r[i, j] if cell(i, j) else 0.0

    S = np.zeros((5, 2, C, NL + 1))
    prev = np.zeros((5, NL, C)); Ep = np.zeros((5, NL, C)); cxW = np.zeros((6, NL, C))
    dC = np.zeros((NL, C)); dW = np.zeros((NL, C))
    xout = x.copy(); rout = r.copy(); dfin = np.zeros_like(r)
    for t in range(1, t_end + 1):
        par, pp = t & 1, (t - 1) & 1
        Snew = S.copy()
        # stage 0
        for L in range(NL):
            i0 = t - L
            _, _, ninv, _ = entry(t, L)
            for c in range(C):
                Snew[0, par, c, L] = rval(i0, L * C + 1 + c) * (-ninv[c])
        # sweeps
        for g in range(1, 5):
            newprev = np.zeros((NL, C))
            for L in range(NL):
                ig = t - L - 2 * g
                cxE, _, _, _ = entry(t - 2 * g + 1, L)
                _, cy, ninv, _ = entry(t - 2 * g, L)
                E = S[g - 1, pp, :, L].copy()
                Nx = S[g - 1, pp, 0, L + 1]
                Sl = prev[g, L - 1, C - 1] if L > 0 else 0.0
                res = np.zeros(C)
                for c in range(C):
                    Sop = Sl if c == 0 else res[c - 1]
                    Nop = Nx if c == C - 1 else Ep[g, L, c + 1]
                    res[c] = (prev[g, L, c] * cxW[g, L, c] + E[c] * cxE[c] + Sop * cy[c] + Nop * cy[c + 1] - rval(ig, L * C + 1 + c)) * ninv[c]
                Snew[g, par, :, L] = res
                newprev[L] = res; Ep[g, L] = E; cxW[g, L] = cxE
            prev[g] = newprev
        # stage 5
        dE = S[4, pp, :, :NL].T.copy()            # (NL, C)
        Nx5 = S[4, pp, 0, 1:NL + 1].copy()
        ndW = dC.copy(); ndC = Ep[0].copy()       # Ep[0] used as "previous dE"
        for L in range(NL):
            i5 = t - L - LAG
            cxE, _, _, _ = entry(t - LAG + 1, L)
            _, cy, _, dg = entry(t - LAG, L)
            Sl = ndW[L - 1, C - 1] if L > 0 else 0.0
            for c in range(C):
                j = L * C + 1 + c
                if not cell(i5, j):
                    continue
                c_ = ndC[L, c]
                w_ = c_ if i5 == 1 else ndW[L, c]
                e_ = c_ if i5 == ni else dE[L, c]
                s_ = c_ if j == 1 else (Sl if c == 0 else ndC[L, c - 1])
                n_ = c_ if j == mj else (Nx5[L] if c == C - 1 else ndC[L, c + 1])
                Ad = c_ * dg[c] + w_ * cxW[5, L, c] + e_ * cxE[c] + s_ * cy[c] + n_ * cy[c + 1]
                rout[i5, j] = r[i5, j] - Ad
                xout[i5, j] = x[i5, j] + c_
                dfin[i5, j] = c_
            cxW[5, L] = cxE
        dW[:] = ndW; dC[:] = ndC                   # registers carried to the next step
        Ep[0] = dE
        S = Snew
    return dfin, xout, rout


def main():
    rng = np.random.default_rng(0)
    for (ni, mj, C) in [(12, 12, 1), (20, 12, 2), (16, 18, 3), (24, 36, 6), (10, 7, 2), (9, 70, 3)]:
        n, m = ni + 2, mj + 2
        r = rng.normal(size=(n, m)); r[0, :] = r[-1, :] = 0; r[:, 0] = r[:, -1] = 0
        lx = rng.uniform(0.1, 0.3, size=(n + 1, m + 1)); ly = rng.uniform(0.1, 0.3, size=(n + 1, m + 1))
        inv = np.ones((n, m)); diag = np.zeros((n, m))
        for i in range(1, n - 1):
            for j in range(1, m - 1):
                s = lx[i, j] + lx[i + 1, j] + ly[i, j] + ly[i, j + 1]
                diag[i, j] = -s; inv[i, j] = -1 / s
        x = rng.normal(size=(n, m))
        d_ref, x_ref, r_ref = reference(r.copy(), lx, ly, inv, diag, x)
        d, xo, ro = pipelined(r, lx, ly, inv, diag, x, C)
        ok = (np.array_equal(d[1:-1, 1:-1], d_ref[1:-1, 1:-1]) and np.array_equal(xo[1:-1, 1:-1], x_ref[1:-1, 1:-1])
              and np.array_equal(ro[1:-1, 1:-1], r_ref[1:-1, 1:-1]))
        print((ni, mj, C), "OK" if ok else "MISMATCH", np.abs(d[1:-1, 1:-1] - d_ref[1:-1, 1:-1]).max())


if __name__ == "__main__":
    main()
