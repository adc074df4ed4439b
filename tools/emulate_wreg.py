"""Numpy emulation of the general register-FFT kernel's index algebra (csrc/kernel_wreg.cuh):
T = M/32 threads per frame, 32 points per thread, radix-2 DIT stages 1-5 / 6-10 / 11-12 in registers
with in-place shared-memory tiles A[row][col] (row stride 33).  Checks against numpy.fft.rfft for
M = 32 ... 4096.  Development aid only."""
import numpy as np


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def dit_stages(a, nst, tw):
    """a: complex array (one sub-FFT, natural order of the local index q);
    tw(u, p): twiddle of local stage u (1..nst), position p < 2^(u-1)."""
    n = 1 << nst
    for u in range(1, nst + 1):
        half = 1 << (u - 1)
        for blk in range(0, n, 2 * half):
            for p in range(half):
                w = tw(u, p)
                x, y = a[blk + p], a[blk + p + half] * w
                a[blk + p], a[blk + p + half] = x + y, x - y


def run(log2m, xw):
    M = 1 << log2m
    N = 2 * M
    T = M // 32
    R = log2m - 5
    R2 = min(R, 5)
    R3 = R - R2
    L2 = 1 << R3
    z = xw[0::2] + 1j * xw[1::2]
    A = np.zeros((T, 32), complex)           # shared tile, row b, col k_a
    # ---- pass 1: thread b, stages 1-5
    for b in range(T):
        a = np.zeros(32, complex)
        for j in range(32):
            a[bitrev(j, 5)] = z[b + T * j]
        dit_stages(a, 5, lambda u, p: np.exp(-2j * np.pi * p / (1 << u)))
        A[b, :] = a
    # ---- pass 2: R2 stages (global stages 6 .. 5+R2)
    G2 = 32 >> R2
    for t in range(T):
        if R <= 5:
            subs = [(t + T * c, 0) for c in range(G2)]          # (k_a, hi2')
        else:
            subs = [(t % 32, t // 32)]
        for (ka, hi2p) in subs:
            a = np.array([A[bitrev(q, R2) * L2 + hi2p, ka] for q in range(1 << R2)])
            dit_stages(a, R2, lambda u, p: np.exp(-2j * np.pi * (p * 32 + ka) / (32 << u)))
            for q in range(1 << R2):
                A[bitrev(q, R2) * L2 + hi2p, ka] = a[q]
    # ---- pass 3: R3 stages (global stages 11 ..)
    if R3:
        G3 = 32 >> R3
        for t in range(T):
            ka, g = t % 32, t // 32
            for c in range(G3):
                q = g * G3 + c
                rows = [bitrev(q, 5) * L2 + bitrev(h, R3) for h in range(1 << R3)]
                a = np.array([A[r, ka] for r in rows])
                dit_stages(a, R3, lambda u, p: np.exp(-2j * np.pi * (p * 1024 + q * 32 + ka) / (1024 << u)))
                for h in range(1 << R3):
                    A[rows[h], ka] = a[h]
    # ---- Z[k] at A[bitrev_R(k >> 5)][k & 31]
    Z = np.fft.fft(z)
    for k in range(M):
        assert np.allclose(A[bitrev(k >> 5, R), k & 31], Z[k]), (log2m, k)
    # ---- untangle: thread t handles k = t + T*i, i < 16
    X = np.zeros(M + 1, complex)
    get = lambda k: A[bitrev((k % M) >> 5, R), (k % M) & 31]
    for t in range(T):
        for i in range(16):
            k = t + T * i
            zk, zm = get(k), get(M - k)
            e, o = zk + np.conj(zm), zk - np.conj(zm)
            tt = o * (-1j) * np.exp(-2j * np.pi * k / N)
            X[k] = 0.5 * (e + tt)
            if k:
                X[M - k] = 0.5 * np.conj(e - tt)
    X[M // 2] = np.conj(get(M // 2))
    ref = np.fft.rfft(xw)
    assert np.allclose(X[:M], ref[:M])
    print(f"M={M}: T={T} R2={R2} R3={R3} ok, max err {np.abs(X[:M] - ref[:M]).max():.2e}")


def merged_pass3(log2m, A, z):
    """The merged last pass of M = 2048 / 4096 (kernel_wreg.cuh, R3 > 0): thread (pi = t & 15, qg = t >> 4) takes the
    column pair (pi, 32 - pi) (pair 0: columns 0 and 16) and the q's whose mirrors it also holds, so every Z[k] meets
    Z[M - k] in the same thread's registers.  A is the tile after pass 2.  Returns the M bins."""
    M = 1 << log2m
    N = 2 * M
    T = M // 32
    R = log2m - 5
    R3 = R - 5
    S3 = 1 << R3
    NP = 16 // (T // 16)       # (q, mirror q) pairs per column per thread
    NQ = 2 * NP
    X = np.full(M + 1, np.nan, complex)
    seen = np.zeros(M + 1, int)

    def untangle(zk, zm, k):
        e, o = zk + np.conj(zm), zk - np.conj(zm)
        tt = o * (-1j) * np.exp(-2j * np.pi * k / N)
        return 0.5 * (e + tt), 0.5 * np.conj(e - tt)

    for t in range(T):
        pi, qg = t & 15, t >> 4
        self_ = pi == 0
        self0 = self_ and qg == 0
        col = [pi, 16 if self_ else 32 - pi]
        q = [[0] * NQ, [0] * NQ]
        for s in range(NP):
            j = qg * NP + s
            if self_:
                q[0][2 * s], q[0][2 * s + 1] = j, (16 if j == 0 else 32 - j)
                q[1][2 * s], q[1][2 * s + 1] = 31 - j, j
            else:
                q[0][2 * s], q[0][2 * s + 1] = j, 31 - j
                q[1][2 * s], q[1][2 * s + 1] = 31 - j, j
        v = np.zeros((2, NQ, S3), complex)
        for x in range(2):
            for i in range(NQ):
                qq, ka = q[x][i], col[x]
                a = np.array([A[bitrev(qq, 5) * S3 + bitrev(h, R3), ka] for h in range(S3)])
                dit_stages(a, R3, lambda u, p: np.exp(-2j * np.pi * (p * 1024 + qq * 32 + ka) / (1024 << u)))
                v[x, i] = a      # v[x, i, h] = Z[col + 32 (q + 32 h)]
        for x in range(2):
            for i in range(NQ):
                for h in range(S3 // 2):
                    k = col[x] + 32 * (q[x][i] + 32 * h)
                    zm = v[1 - x, i, S3 - 1 - h]
                    if self_:
                        zm = v[x, i ^ 1, S3 - 1 - h]
                    if self0 and x == 0 and i == 0:
                        zm = v[0, 0, (S3 - h) % S3]
                    if self0 and x == 0 and i == 1:
                        zm = v[0, 1, S3 - 1 - h]
                    xk, xm = untangle(v[x, i, h], zm, k)
                    mk = M - k
                    if k == 0:
                        xm, mk = np.conj(v[0, 0, S3 // 2]), M // 2
                    X[k], X[mk] = xk, xm
                    seen[k] += 1
                    seen[mk] += 1
    assert (seen[:M] == 1).all() and seen[M] == 0, np.nonzero(seen[:M] != 1)
    return X[:M]


def run_merged(log2m, xw):
    """passes 1-2 as run(), then the merged pass 3"""
    M = 1 << log2m
    T = M // 32
    R = log2m - 5
    R3 = R - 5
    L2 = 1 << R3
    z = xw[0::2] + 1j * xw[1::2]
    A = np.zeros((T, 32), complex)
    for b in range(T):
        a = np.zeros(32, complex)
        for j in range(32):
            a[bitrev(j, 5)] = z[b + T * j]
        dit_stages(a, 5, lambda u, p: np.exp(-2j * np.pi * p / (1 << u)))
        A[b, :] = a
    for t in range(T):
        ka, hi2p = t % 32, t // 32
        a = np.array([A[bitrev(q, 5) * L2 + hi2p, ka] for q in range(32)])
        dit_stages(a, 5, lambda u, p: np.exp(-2j * np.pi * (p * 32 + ka) / (32 << u)))
        for q in range(32):
            A[bitrev(q, 5) * L2 + hi2p, ka] = a[q]
    X = merged_pass3(log2m, A, z)
    ref = np.fft.rfft(xw)
    assert np.allclose(X, ref[:M])
    print(f"M={M}: merged pass 3 ok, max err {np.abs(X - ref[:M]).max():.2e}")


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for lm in range(5, 13):
        run(lm, rng.standard_normal(2 << lm))
    for lm in (11, 12):
        run_merged(lm, rng.standard_normal(2 << lm))
