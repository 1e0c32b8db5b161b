"""CPU: the C-ABI library loads, exports every symbol include/sciport_b200.h declares, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sciport_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from sciport_rs_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
    assert set(_lib.EXPORTS) == set(names)


def test_version_and_argument_checks():
    from sciport_rs_b200 import _lib
    assert _lib.lib.spb_version() >= 100
    assert _lib.lib.spb_device_count() >= 0


def test_no_cpu_fallback_without_device():
    import torch
    from sciport_rs_b200 import _lib, special
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.SpbError) as e:
        special.jv(2.5, np.array([1 + 1j]))
    assert e.value.code == _lib.E_NO_DEVICE
    with pytest.raises(_lib.SpbError):
        special.gamma(np.array([1.5]))
    with pytest.raises(_lib.SpbError):
        special.i0(np.array([0.5]))


def test_product_does_not_import_oracle():
    """The package must not reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "sciport_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert not re.search(r"#include\s+\"[^\"]*oracle/", text), f
                assert "libspb_oracle" not in text, f


def test_header_enums_match_the_python_binding():
    """exec.flags and the kernel ids of spb_last_profile are duplicated in sciport_rs_b200/_lib.py: keep them in step
    with include/sciport_b200.h."""
    from sciport_rs_b200 import _lib
    src = open(os.path.join(ROOT, "include", "sciport_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    enum = {k: int(v) for k, v in re.findall(r"\b(SPB_[A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    assert enum["SPB_SEM_SCIPY"] == _lib.SEM_SCIPY and enum["SPB_RAW_NEG_ORDERS"] == _lib.RAW_NEG_ORDERS
    assert enum["SPB_PATH_MONOLITHIC"] == _lib.PATH_MONOLITHIC and enum["SPB_PROFILE"] == _lib.PROFILE
    assert enum["SPB_PATH_NO_DIRECT"] == _lib.PATH_NO_DIRECT
    flags = [enum[k] for k in ("SPB_SEM_SCIPY", "SPB_RAW_NEG_ORDERS", "SPB_PATH_MONOLITHIC", "SPB_PROFILE", "SPB_PATH_NO_DIRECT")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags)  # distinct single bits
    ids = {k[len("SPB_K_"):].lower(): v for k, v in enum.items() if k.startswith("SPB_K_")}
    assert sorted(ids.values()) == list(range(len(_lib.KERNEL_NAMES)))
    for name, i in ids.items():
        want = "direct_asyi" if name == "direct" else name
        assert _lib.KERNEL_NAMES[i] == want, (name, i, _lib.KERNEL_NAMES[i])
    assert enum["SPB_E_NO_DEVICE"] == _lib.E_NO_DEVICE and enum["SPB_E_ARG"] == _lib.E_ARG
