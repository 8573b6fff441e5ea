"""CPU: the exact constant-division helper of the CUDA path (ev2gym_b200/csrc/ev2b_math.h) equals IEEE `/`.

The device code replaces `x / d` by q = x*rd; r = fma(-d, q, x); fma(rd, r, q) with rd = RN(1/d) for
divisors fixed at load time (1000, 100, 1e5, 60, the timescale, each EV model's battery capacity,
each charger's effective voltage).  Bit-exactness of the battery level rests on this identity, so it
is checked here on 2*10^6 operands per divisor (lattice values k/100, random mantissas, uniform)."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ev2gym_b200", "csrc")


def test_div_const_matches_ieee_division():
    exe = os.path.join(tempfile.mkdtemp(), "div_const_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(CSRC, "div_const_check.c"), "-lm"])
    volts = [v * np.sqrt(k) for v in (230.0, 400.0) for k in (1, 2, 3)]
    caps = [28.5, 40.0, 50.0, 57.5, 64.0, 77.0, 82.0, 100.0, 16.7, 39.2]
    divisors = [1000, 100, 100000, 60, 15, 10, 5, 7, 4, 0.25, 6, 2.4] + caps + volts
    out = subprocess.check_output([exe, "2000000"] + [repr(float(d)) for d in divisors], text=True)
    rows = [l.split() for l in out.strip().splitlines()]
    assert len(rows) == len(divisors) and all(int(r[1]) == 0 for r in rows), rows


def test_quotient_at_least_one_is_a_plain_comparison():
    """ev_step_item replaces the reference's `1 <= (pts - soc) / pilot` (ev.py:323) by `pilot <= pts - soc`: for positive
    float64 operands the rounded quotient is >= 1 exactly when x >= y (x < y gives x / y < 1 - 2^-53, which rounds
    below 1).  Checked on operands a few ulps around equality and on random ratios, and for a zero numerator of the
    saturation test `(pilot - maxd) / maxd` (0 / d * (ts - 1) must leave ts unchanged)."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    y = np.exp(rng.uniform(-12, 3, n))
    k = rng.integers(-4, 5, n)
    x = y.copy()
    for _ in range(4):
        x = np.where(k > 0, np.nextafter(x, np.inf), np.where(k < 0, np.nextafter(x, -np.inf), x))
        k = k - np.sign(k)
    for xx in (x, y * rng.uniform(0.5, 1.5, n)):
        assert np.array_equal(1.0 <= xx / y, y <= xx)
    ts = rng.uniform(0.0, 1.0, n).round(3)
    assert np.array_equal(ts + (0.0 / y) * (ts - 1.0), ts)
