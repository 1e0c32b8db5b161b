import os
import sys

import pytest

# The direct large-|z| kernel of broadcast-order jobs (k_direct.cu) is launched for chunks of 2^22 points and more; the
# parity tests use far smaller batches, so the threshold is lowered for the test process (read once by the library,
# on the first call): every broadcast-order test then runs through the kernel, and path="classified" is the other side.
os.environ.setdefault("SPB_DIRECT_MIN_LOG2", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
