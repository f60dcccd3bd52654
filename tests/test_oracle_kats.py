"""Pins the oracle's arithmetic against the only known-answer tests the reference holds for this path
(stirling_numbers/src/lib.rs:259-318) and against exact rational / mpmath evaluation."""
from fractions import Fraction

import mpmath as mp
import numpy as np
import pytest

from oracle import pyoracle


def exact_stirling2_row(n):
    row = [1]  # S(0,0)
    for m in range(1, n + 1):
        new = [0] * (m + 1)
        for k in range(1, m + 1):
            new[k] = (row[k] * k if k < m else 0) + row[k - 1]
        row = new
    return row


def sr_exact(n, k, s2):
    return Fraction(s2[k] * mp_fact(k), k ** n)


def mp_fact(k):
    f = 1
    for i in range(2, k + 1):
        f *= i
    return f


def test_stirling_10_5_is_42525():
    # stirling_numbers/src/lib.rs:261: S2(10,5) == 42525; sr[n][k] = S(n,k) k!/k^n
    hi, lo = pyoracle.sr_entry(10, 5)
    v = (Fraction(hi) + Fraction(lo)) * Fraction(5 ** 10, 120)
    assert abs(v - 42525) < Fraction(1, 10 ** 24)


@pytest.mark.parametrize("n", [219, 722])
def test_ratio_table_digits(n):
    # lib.rs:271-306 demands 14 digits at n=219 (table) and 12 digits at n=722 (ratio table) from f64;
    # the double-double table must do far better wherever the value is a normal f64
    s2 = exact_stirling2_row(n)
    worst = 0
    for k in range(1, n + 1):
        hi, lo = pyoracle.sr_entry(n, k)
        ex = sr_exact(n, k, s2)
        if ex < Fraction(1, 10 ** 280):
            continue  # n!/n^n underflows to denormals (722!/722^722 ~ 1e-312), as in the reference
        rel = abs((Fraction(hi) + Fraction(lo)) - ex) / ex
        worst = max(worst, rel)
    assert worst < Fraction(1, 10 ** 26)


def test_p_27_30_2500_formats_as_0_0005953():
    # lib.rs:310-315
    assert f"{pyoracle.p1(30, 3, 2500):.7f}" == "0.0005953"


def p1_exact(k, d, n):
    s2 = exact_stirling2_row(k)
    p = Fraction(1)
    for u in range(k - d + 1, k + 1):
        z = Fraction(s2[u] * mp_fact(u), u ** k) * Fraction(u, n) ** k
        for v in range(1, u + 1):
            z *= Fraction(n - v + 1, u - v + 1)
        p -= z
    return max(p, Fraction(0))


@pytest.mark.parametrize("k,d,n", [(8, 2, 2445), (20, 5, 2445), (30, 3, 2500), (16, 3, 2076), (40, 10, 2445), (12, 1, 1800)])
def test_p1_against_exact_rational(k, d, n):
    got = pyoracle.p1(k, d, n)
    ex = p1_exact(k, d, n)
    # values well above double-double resolution relative to 1 (~1e-32) must be correctly rounded-ish
    if ex > Fraction(1, 10 ** 16):
        assert abs(Fraction(got) - ex) / ex < Fraction(1, 10 ** 12)
    else:
        assert got >= 0.0


def test_p1_edge_cases():
    assert pyoracle.p1(0, 0, 2000) == 1.0   # no draws: empty sum
    assert pyoracle.p1(5, 0, 2000) == 1.0   # d = 0: empty sum
    with pytest.raises(ValueError):
        pyoracle.p1(3001, 1, 9000)          # the reference indexes sr[3001] and panics


def test_dd_ops_against_mpmath():
    mp.mp.dps = 70
    rng = np.random.default_rng(3)
    for _ in range(200):
        a = (float(rng.integers(1, 4000)), 0.0)
        b = (float(rng.integers(1, 4000)), 0.0)
        q = pyoracle.dd_ops(a, b)[3]
        qa = (q[0], q[1])
        r = pyoracle.dd_ops(qa, (float(rng.integers(1, 4000)) / 7.0, 0.0))
        ma, mb = mp.mpf(qa[0]) + mp.mpf(qa[1]), None
        x = float(rng.integers(1, 4000)) / 7.0
        # recompute with the same operand to compare
        r = pyoracle.dd_ops(qa, (x, 0.0))
        mb = mp.mpf(x)
        want = [ma + mb, ma - mb, ma * mb, ma / mb]
        for (hi, lo), w in zip(r, want):
            got = mp.mpf(hi) + mp.mpf(lo)
            assert abs(got - w) <= abs(w) * mp.mpf(10) ** -30 + mp.mpf(10) ** -300
