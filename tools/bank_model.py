"""Shared-memory bank-conflict model for the dim-1 kernels' access patterns, and a search for the
XOR swizzle  phys(i) = i ^ f(i)  (f = GF(2)-linear map of the high index bits into the low 4 bits,
i.e. into the 16 eight-byte bank pairs) that makes every FFT pass conflict-free.

float2 (64-bit) shared accesses are served per half-warp: 16 lanes x 8 B = 128 B; a half-warp
costs as many wavefronts as the most-loaded bank pair has distinct addresses.
"""
import itertools
import sys

import numpy as np

PLAN = {32: (8, 4), 64: (8, 8), 128: (16, 8), 256: (16, 16), 512: (8, 8, 8), 1024: (16, 8, 8), 2048: (16, 16, 8),
        4096: (16, 16, 16)}


def cfg(LM):
    M = 1 << LM
    NT = M if LM <= 8 else (256 if LM <= 11 else 512)
    RPT = M // NT
    return M, NT, RPT


def stage_params(M, s):
    R = PLAN[M][s]
    LS = M
    for q in range(s):
        LS //= PLAN[M][q]
    return R, LS, LS // R


def patterns(LM):
    """Yields (name, array[n_requests, 32] of float2 indices c*M + i) for one block."""
    M, NT, RPT = cfg(LM)
    plan = PLAN[M]
    NS = len(plan)
    tids = np.arange(NT)
    out = []

    def items(ITEMS):
        # (wi, c) of each thread for the first loop iteration
        if NT >= ITEMS:
            return tids % ITEMS, tids // ITEMS
        return tids, np.zeros_like(tids)   # wi = tid (first trip), c = 0

    for s in range(NS):
        R, LS, ST = stage_params(M, s)
        ITEMS = M // R
        wi, c = items(ITEMS)
        base = (wi // ST) * LS + (wi % ST)
        reqs = np.stack([c * M + base + m * ST for m in range(R)])     # [R, NT]
        out.append((f"stage{s}", reqs.reshape(R * (NT // 32), 32) if False else reqs.reshape(R, NT // 32, 32).reshape(-1, 32)))
    # stencil: rows i0 = tid*RPT .. ; reads at row offsets -1..RPT of a column
    i0 = tids * RPT
    reqs = np.stack([(i0 + r) % M for r in range(-1, RPT + 1)])
    out.append(("stencil", reqs.reshape(-1, NT // 32, 32).reshape(-1, 32)))
    return out


def wavefronts(addr, f):
    """addr [n,32] logical indices; f maps index array -> physical index array."""
    phys = f(addr)
    tot = 0
    for half in (phys[:, :16], phys[:, 16:]):
        bank = half & 15
        # distinct addresses per bank pair
        w = np.zeros(half.shape[0], dtype=np.int64)
        for b in range(16):
            sel = bank == b
            # count distinct addresses among selected lanes per row
            vals = np.where(sel, half, -1)
            vals.sort(axis=1)
            distinct = ((vals[:, 1:] != vals[:, :-1]) & (vals[:, 1:] >= 0)).sum(axis=1) + (vals[:, 0] >= 0)
            w = np.maximum(w, distinct)
        tot += int(w.sum())
    return tot


def make_f(LM, A):
    """A: list of 4 ints; bit r of the low nibble is XORed with parity(i_high & A[r]), i_high = i >> 4."""
    M = 1 << LM

    def f(idx):
        i = idx % M
        hi = i >> 4
        x = np.zeros_like(i)
        for r in range(4):
            p = hi & A[r]
            # parity
            p ^= p >> 8; p ^= p >> 4; p ^= p >> 2; p ^= p >> 1
            x |= (p & 1) << r
        return idx ^ x
    return f


def ideal(addr):
    return addr.shape[0] * 2


def search(LM, iters=4000, seed=0):
    rng = np.random.default_rng(seed)
    pats = patterns(LM)
    nb = LM - 4

    def cost(A):
        f = make_f(LM, A)
        return sum(wavefronts(a, f) for _, a in pats)

    best = [0, 0, 0, 0]
    bc = cost(best)
    base = bc
    lo = sum(ideal(a) for _, a in pats)
    for it in range(iters):
        A = list(best)
        r = rng.integers(4)
        A[r] ^= 1 << int(rng.integers(nb))
        if rng.random() < 0.3:
            r2 = rng.integers(4)
            A[r2] ^= 1 << int(rng.integers(nb))
        c = cost(A)
        if c <= bc:
            best, bc = A, c
            if bc == lo:
                break
    return best, bc, base, lo


if __name__ == "__main__":
    for LM in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else range(5, 13):
        A, c, base, lo = search(LM)
        f = make_f(LM, A)
        detail = {n: (wavefronts(a, f), ideal(a)) for n, a in patterns(LM)}
        print(f"LM={LM} A={[hex(a) for a in A]} cost {c} (identity {base}, ideal {lo}) {detail}")


# ---- restricted family: f(i, c) = ((i >> s1) & m1) ^ ((i >> s2) & m2) ^ ((c << t) & 15) -----------------
def make_f2(LM, s1, m1, s2, m2, t):
    M = 1 << LM

    def f(idx):
        i = idx % M
        c = idx // M
        x = ((i >> s1) & m1) ^ ((i >> s2) & m2)
        if t >= 0:
            x ^= (c << t) & 15
        return idx ^ x
    return f


def search2(LM):
    pats = patterns(LM)
    lo = sum(ideal(a) for _, a in pats)
    best = None
    for s1 in range(2, LM):
        for m1 in (15, 7, 14, 12, 8, 3, 6, 1, 2, 4):
            for s2 in [0] + list(range(s1 + 1, LM)):
                for m2 in ((0,) if s2 == 0 else (15, 8, 12, 14, 7, 3, 1, 4, 2, 6)):
                    for t in (-1, 0, 1, 2, 3):
                        f = make_f2(LM, s1, m1, s2 if s2 else 31, m2, t)
                        c = sum(wavefronts(a, f) for _, a in pats)
                        key = (c, (s2 != 0) + (t >= 0))
                        if best is None or key < best[0]:
                            best = (key, (s1, m1, s2, m2, t))
    (c, _), prm = best
    f = make_f2(LM, prm[0], prm[1], prm[2] if prm[2] else 31, prm[3], prm[4])
    detail = {n: (wavefronts(a, f), ideal(a)) for n, a in pats}
    return prm, c, lo, detail


if __name__ == "__main__" and "--family2" in sys.argv:
    pass
