"""Numpy emulation of the warp-per-frame 32x32 kernel's index algebra (spectrogram_b200/csrc/
kernel_w32.cuh).  Lanes are axis 0, per-lane registers axis 1.  Run: python tools/emulate_w32.py
Checks the register FFT, exchange layout, fused pass-2 twiddles and the real-FFT untangle against
numpy.fft.rfft.  Development aid only (not imported by the package or the tests)."""
import numpy as np

M, L, P = 1024, 32, 32
N = 2 * M


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def fft32_dit_const(a):
    """a: [lanes, 32] holding bit-reversed input; in place radix-2 DIT, natural output."""
    for s in range(1, 6):
        half = 1 << (s - 1)
        for blk in range(0, 32, 2 * half):
            for p in range(half):
                w = np.exp(-2j * np.pi * p / (2 * half))
                x, y = a[:, blk + p].copy(), a[:, blk + p + half] * w
                a[:, blk + p], a[:, blk + p + half] = x + y, x - y


def tables():
    lane = np.arange(L)
    tw2 = np.zeros((31, L), complex)
    for us in range(1, 6):
        half = 1 << (us - 1)
        for p in range(half):
            tw2[half - 1 + p] = np.exp(-2j * np.pi * (p * 32 + lane) / (32 * 2 * half))
    ut = np.zeros((16, L), complex)
    for i in range(16):
        ut[i] = np.exp(-2j * np.pi * (lane + 32 * i) / N)
    return tw2, ut


def run(xw):
    tw2, ut = tables()
    lane = np.arange(L)
    z = xw[0::2] + 1j * xw[1::2]
    a = np.zeros((L, P), complex)
    for j in range(32):
        a[:, bitrev(j, 5)] = z[lane + 32 * j]
    fft32_dit_const(a)                       # a[lane=b][k_a]
    xbuf = np.zeros(32 * 34, complex)
    for ka in range(32):
        xbuf[lane * 34 + ka] = a[:, ka]
    u = np.zeros((L, P), complex)            # lane = k_a, reg q2
    for q2 in range(32):
        u[:, q2] = xbuf[bitrev(q2, 5) * 34 + lane]
    for us in range(1, 6):
        half = 1 << (us - 1)
        for blk in range(0, 32, 2 * half):
            for p in range(half):
                w = tw2[half - 1 + p]
                x, y = u[:, blk + p].copy(), u[:, blk + p + half] * w
                u[:, blk + p], u[:, blk + p + half] = x + y, x - y
    # u[lane][kb] = Z[lane + 32 kb]
    Z = np.fft.fft(z)
    for kb in range(32):
        assert np.allclose(u[:, kb], Z[lane + 32 * kb]), kb
    # untangle exchange: upper half to smem, index k-512; lane 0 adds Z[0] at 512
    ub = np.zeros(513, complex)
    for i in range(16, 32):
        ub[lane + 32 * (i - 16)] = u[:, i]
    ub[512] = u[0, 0]
    X = np.zeros(M + 1, complex)
    for i in range(16):
        k = lane + 32 * i
        zk = u[:, i]
        zm = ub[512 - lane - 32 * i]
        e = zk + np.conj(zm)                 # 2E
        o = zk - np.conj(zm)                 # 2i O
        t = o * (-1j) * ut[i]                # 2 W O
        xk = 0.5 * (e + t)
        xm = 0.5 * np.conj(e - t)
        X[k] = xk
        mk = M - k
        X[mk] = xm                            # (lane 0, i 0) -> index 1024 = Nyquist, discarded
    X[512] = np.conj(u[0, 16])               # lane 0 special: bin 512 from its own register
    ref = np.fft.rfft(xw)
    assert np.allclose(X[:M], ref[:M]), np.abs(X[:M] - ref[:M]).max()
    print("w32 emulation OK, max err", np.abs(X[:M] - ref[:M]).max())


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    run(rng.standard_normal(N))
