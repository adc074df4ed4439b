"""Tolerances of the parity tests (BASELINE.json north_star: 1e-4 relative in float magnitude,
1e-3 dB, +-1 LSB on bytes), made precise for a float32 transform:

  magnitude  |d| <= 1e-4*|ref| + 3.5e-7*frame_peak
             (a float32 FFT has an absolute error floor of ~2-3e-7 of the frame's largest bin; a
             purely relative bound cannot hold on bins ~140 dB below the peak)
  dB         |d| <= 1e-3 dB on bins within 60 dB of the frame peak; elsewhere the bound the
             magnitude tolerance implies
  byte       |d| <= 1 LSB everywhere, at most 0.2 % of the bytes off by that one level
The reference against which these are taken is the float64 oracle (oracle/analyser_oracle.py).

How much of them the kernels use is measured by tools/tolerance_report.py (profiles/r02_tolerance_report.txt: 20
kernel/shape cases x 4 signals).  Worst cases over every kernel family: |d|/peak 3.4e-7; the magnitude bound is 82 %
used with a floor of 2e-7 and exceeded (1.64x) with SURVEY section 7's 1e-7, so the floor sits at 3.5e-7 (about 2x
the need); dB error 2.1e-4 within 50 dB of the peak, 7.2e-4 within 60 dB, 2.9e-3 within 70 dB, 7.7e-3 within 80 dB,
so 1e-3 dB cannot be promised at SURVEY's 80 dB with float32 arithmetic and the window sits at 60 dB; byte mismatch
rate at most 6.5e-4, never more than one level.
"""
import numpy as np

MAG_REL = 1e-4
MAG_FLOOR = 3.5e-7
DB_TOL = 1e-3
DB_WINDOW = 60.0


def assert_mag_close(got, ref_mag):
    ref = np.asarray(ref_mag, np.float64)
    got = np.asarray(got, np.float64)
    peak = ref.max(axis=-1, keepdims=True)
    bound = MAG_REL * np.abs(ref) + MAG_FLOOR * peak
    err = np.abs(got - ref)
    bad = err > bound
    assert not bad.any(), f"{bad.sum()} magnitude bins out of tolerance, worst ratio {np.max(err / np.maximum(bound, 1e-300)):.3g}"


def assert_db_close(got_db, ref_mag):
    ref = np.asarray(ref_mag, np.float64)
    got = np.asarray(got_db, np.float64)
    with np.errstate(divide="ignore"):
        ref_db = 20.0 * np.log10(ref)
    peak = ref.max(axis=-1, keepdims=True)
    zero = ref == 0.0
    assert np.all(np.isneginf(got[zero]) | (got[zero] < -300.0)), "zero magnitude must map to -inf dB"
    with np.errstate(divide="ignore", invalid="ignore"):
        near = (~zero) & (ref >= peak * 10 ** (-DB_WINDOW / 20.0))
        err = np.abs(got - ref_db)
        assert np.all(err[near] <= DB_TOL), f"dB error {err[near].max():.3g} > {DB_TOL} within {DB_WINDOW} dB of the peak"
        far = (~zero) & ~near
        rel = MAG_REL + MAG_FLOOR * np.broadcast_to(peak, ref.shape)[far] / ref[far]
        bound = np.where(rel < 0.5, -20.0 * np.log10(np.maximum(1.0 - rel, 1e-12)), np.inf) + DB_TOL
        assert np.all((err[far] <= bound) | ~np.isfinite(bound)), "dB error beyond the magnitude-implied bound"


def assert_bytes_close(got, ref, max_mismatch_frac=0.002):
    d = np.abs(np.asarray(got, np.int32) - np.asarray(ref, np.int32))
    assert d.max(initial=0) <= 1, f"byte differs by {d.max()} LSB"
    assert (d != 0).mean() <= max_mismatch_frac if d.size else True, f"{(d != 0).mean():.3%} of bytes differ by 1 LSB"
