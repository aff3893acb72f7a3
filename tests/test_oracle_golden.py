"""The oracle against the reference's only known-answer vectors on this path:
ring/test_data/test_pol_60_* and test_pol_NTT_60_* (consumed by
ring/ntt_test.go:101-142).  Unlike the Go test (which only compares the first
two coefficients of each limb, ntt_test.go:120-121) every coefficient is checked.
"""
import os

import numpy as np
import pytest

from oracle import ring_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ring_test_data")
SIZES = [8, 16, 32, 64, 128, 256, 512]


def load(name):
    """file format (ntt_test.go:39-99): line0 N, line1 moduli, then one line per limb"""
    with open(os.path.join(GOLD, name)) as f:
        lines = [l for l in f.read().split("\n") if l.strip()]
    N = int(lines[0])
    moduli = [int(x) for x in lines[1].split()]
    coeffs = np.array([[int(x) for x in lines[2 + i].split()] for i in range(len(moduli))], dtype=np.uint64)
    assert coeffs.shape == (len(moduli), N)
    return N, moduli, coeffs


def names(n):
    w = str(n).rjust(4, "_")
    return "test_pol_60_%s_2" % w, "test_pol_NTT_60_%s_2" % w


@pytest.mark.parametrize("n", SIZES)
def test_ntt_golden(n):
    a, b = names(n)
    N, moduli, x = load(a)
    N2, moduli2, want = load(b)
    assert (N, moduli) == (N2, moduli2) and N == n
    ctx = orc.Context(N, moduli)
    got = ctx.ntt(x)
    assert np.array_equal(got, want)
    # InvNTT pinned by round trip (ntt_test.go:125-134)
    back = ctx.invntt(want)
    assert np.array_equal(back, x)


def test_golden_primes_roots():
    # generators found for the two golden primes
    assert orc.lib().orc_primitive_root(576460752303439873) == 15
    assert orc.lib().orc_primitive_root(576460752303702017) == 3


def test_small_primes_table():
    # ring/utils.go:290-391 is exactly the first 2000 primes, ending 17389
    L = orc.lib()
    assert L.orc_small_prime(0) == 2 and L.orc_small_prime(1999) == 17389
    sieve = [p for p in range(2, 17390) if all(p % d for d in range(2, int(p ** 0.5) + 1))]
    assert len(sieve) == 2000
    assert [L.orc_small_prime(i) for i in range(2000)] == sieve
