#!/usr/bin/env python
"""Golden vectors for the Bessel-filter prototype (poles and gain for orders 0..49, three normalisations) from
scipy.signal.besselap — the oracle of /root/reference/tests/signal/bessel.rs:57-71.
Run:  python tests/golden/make_golden_besselap.py"""
import os

import numpy as np
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for norm in ("phase", "delay", "mag"):
    for n in range(0, 50):
        z, p, k = scipy.signal.besselap(n, norm=norm)
        out[f"p_{norm}_{n}"] = np.asarray(p, dtype=complex)
        out[f"k_{norm}_{n}"] = np.float64(k)
np.savez_compressed(os.path.join(HERE, "besselap_golden.npz"), **out)
print(len(out))
