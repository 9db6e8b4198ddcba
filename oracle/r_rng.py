"""oracle/r_rng.py -- TEST INFRASTRUCTURE.

Emulation of the pieces of R's default RNG that the reference's snapshot tests use
(`set.seed(1234); rnorm(n)`, `sample(n, k)`), so that the snapshot values in
/root/reference/tests/testthat/_snaps/kendall-tau.md become usable as golden vectors
without R:

  * set.seed(seed)        : R src/main/RNG.c  (initial scrambling 50x LCG 69069, then
                            625 LCG draws into the Mersenne-Twister seed block,
                            FixupSeeds sets mti = 624)
  * unif_rand()           : MT19937 genrand * 2.3283064365386963e-10, fixup to (0,1)
  * norm_rand(), INVERSION: u = unif_rand(); u = (int)(2^27 * u) + unif_rand();
                            qnorm(u / 2^27)
  * sample(n, k)          : R >= 3.6 "Rejection" sampling (R_unif_index / rbits),
                            partial Fisher-Yates of do_sample

qnorm is evaluated with scipy.special.ndtri instead of R's AS241; both are accurate
to ~1e-16 relative, so the variates can differ from R's by a few ulp.  Everything the
golden tests derive from them depends only on the ORDER of tie-free variates, which
such differences do not change.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import ndtri


class RRng:
    def __init__(self, seed: int):
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
            block = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
                block[j] = s
        key = block[1:].copy()  # dummy[0] = mti is overwritten with 624 by FixupSeeds
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = key
        st["state"]["pos"] = 624
        self._bg.state = st

    def unif_rand(self, size=None):
        raw = self._bg.random_raw(size)
        u = np.asarray(raw, dtype=np.float64) * 2.3283064365386963e-10
        lo = 2.328306437080797e-10
        u = np.where(u <= 0.0, 0.5 * lo, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * lo, u)
        return float(u) if size is None else u

    def rnorm(self, n: int):
        u = self.unif_rand(2 * n)
        u1, u2 = u[0::2], u[1::2]
        big = 134217728.0
        v = np.floor(big * u1) + u2
        return ndtri(v / big)

    def _rbits(self, bits: int) -> float:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return float(v)

    def unif_index(self, dn: float) -> int:
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dn > dv:
                return int(dv)

    def sample(self, n: int, k: int):
        """sample(n, k) without replacement; returns 1-based indices like R."""
        x = list(range(n))
        out = []
        nn = n
        for _ in range(k):
            j = self.unif_index(nn)
            out.append(x[j] + 1)
            nn -= 1
            x[j] = x[nn]
        return np.array(out, dtype=np.int64)
