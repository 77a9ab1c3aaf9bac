"""CPU: gsc_log_cr (soundchunks_b200/csrc/gsc_log.h, shared by the kernels and the oracle) is the correctly rounded
natural logarithm: compared with 300-bit mpmath, which is independent of both."""
import math

import mpmath as mp
import numpy as np


def test_log_cr_is_correctly_rounded(oracle):
    L = oracle.lib()
    mp.mp.prec = 300
    rng = np.random.default_rng(1)
    xs = (list(np.exp(rng.uniform(-700, 700, 3000))) + list(1 + rng.uniform(-1e-3, 1e-3, 2000)) +
          list(rng.uniform(0.5, 2, 3000)) + list(rng.uniform(1e-12, 64.0, 4000)) +     # the range the features use
          [1.0, 2.0, 0.5, 1e-12, 1.0000000000000002, 0.9999999999999999, 5e-324, 1e-310, 1.7976931348623157e308,
           1.4142135623730951, 1.414213562373095, 0.7071067811865476, 0.7071067811865475])
    glibc_off = 0
    for x in xs:
        x = float(x)
        want = float(mp.log(mp.mpf(x)))          # mpmath rounds to the nearest double
        assert L.gsc_ref_log_cr(x) == want, x
        glibc_off += math.log(x) != want
    # (glibc's log is a "< 1 ulp" function: it misses the correctly rounded value now and then, which is why the
    # features used to depend on the platform)
    assert glibc_off < len(xs) // 100


def test_log_cr_special_values(oracle):
    L = oracle.lib()
    assert L.gsc_ref_log_cr(0.0) == -math.inf and L.gsc_ref_log_cr(math.inf) == math.inf
    assert math.isnan(L.gsc_ref_log_cr(-1.0)) and math.isnan(L.gsc_ref_log_cr(math.nan))
    assert L.gsc_ref_log_cr(1.0) == 0.0
