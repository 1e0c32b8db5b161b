#!/usr/bin/env python
"""Golden vectors for the window functions on the special-function path, from scipy.signal.windows — the oracle of
the reference's own tests (/root/reference/tests/signal/fir_filter_design_windows.rs:243-281, tolerance 1e-14;
len < 1000, beta in [2, 30)).  Run:  python tests/golden/make_golden_windows.py"""
import os

import numpy as np
import scipy.signal.windows as w

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260110)
    out = {}
    cases = []
    for i in range(40):
        m = int(rng.integers(2, 1000))
        beta = float(rng.uniform(2.0, 30.0))
        sym = bool(rng.integers(0, 2))
        cases.append((m, beta, sym))
    cases += [(0, 5.0, True), (1, 5.0, True), (2, 5.0, False), (3, 0.0, True), (5000, 14.0, True), (4097, 8.6, False)]
    out["cases"] = np.array([(m, b, s) for m, b, s in cases], dtype=np.float64)
    for i, (m, beta, sym) in enumerate(cases):
        out[f"kaiser_{i}"] = w.kaiser(m, beta, sym=sym)
        out[f"lanczos_{i}"] = w.lanczos(m, sym=sym)
        me = m + (m % 2)
        out[f"kbd_{i}"] = w.kaiser_bessel_derived(me, beta, sym=True) if me >= 2 else np.empty(0)
    np.savez_compressed(os.path.join(HERE, "window_golden.npz"), **out)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
