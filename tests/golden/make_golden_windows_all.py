#!/usr/bin/env python
"""Golden vectors for the elementwise window functions from scipy.signal.windows, the oracle of the
reference's own tests (/root/reference/tests/signal/fir_filter_design_windows.rs:29-320: len < 1000, random sym, the
parameter ranges below).  Run:  python tests/golden/make_golden_windows_all.py"""
import os

import numpy as np
import scipy.signal.windows as w

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["boxcar", "triang", "blackman", "hamming", "hann", "bartlett", "flattop", "parzen", "bohman",
         "blackmanharris", "nuttall", "barthann", "cosine", "exponential", "tukey", "taylor", "gaussian",
         "general_gaussian", "general_cosine", "general_hamming"]


def main():
    rng = np.random.default_rng(20261020)
    out = {}
    cases = []
    for i in range(24):
        m = int(rng.integers(1, 1000))
        cases.append((m, bool(rng.integers(0, 2)), float(rng.uniform(0.0, 1.0)), float(rng.uniform(2.0, 30.0)),
                      float(rng.uniform(0.0, 1.0))))
    cases += [(0, True, .5, 5., .5), (1, False, .5, 5., .5), (2, False, .5, 5., .5), (2, True, .3, 3., .7),
              (3, True, .9, 2., .1), (4096, False, .25, 17., .6), (5001, True, .75, 29., .9)]
    out["cases"] = np.array(cases, dtype=np.float64)  # m, sym, alpha (tukey / general_hamming / power), sigma, p
    for i, (m, sym, alpha, sig, p) in enumerate(cases):
        for name in NAMES[:13]:
            out[f"{name}_{i}"] = getattr(w, name)(m, sym=sym)
        out[f"exponential_{i}"] = w.exponential(m, sym=sym)  # reference test: center None, tau None (= 1)
        out[f"exponential_tau_{i}"] = w.exponential(m, center=None if sym else 0.3 * m, tau=sig, sym=sym) if not sym \
            else w.exponential(m, tau=sig, sym=True)
        out[f"tukey_{i}"] = w.tukey(m, alpha=alpha, sym=sym)
        out[f"taylor_{i}"] = w.taylor(m, norm=False, sym=sym)
        out[f"taylor_n6_{i}"] = w.taylor(m, nbar=6, sll=45.0, norm=True, sym=sym)
        out[f"gaussian_{i}"] = w.gaussian(m, sig, sym=sym)
        out[f"general_gaussian_{i}"] = w.general_gaussian(m, p, sig, sym=sym)
        out[f"general_cosine_{i}"] = w.general_cosine(m, [0.3, 0.4, 0.2, 0.07, 0.02, 0.01], sym=sym)
        out[f"general_hamming_{i}"] = w.general_hamming(m, alpha, sym=sym)
    np.savez_compressed(os.path.join(HERE, "window_all_golden.npz"), **out)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
