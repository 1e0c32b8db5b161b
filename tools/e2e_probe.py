#!/usr/bin/env python
"""End-to-end (host buffers in, host buffers out) throughput of special.ive_kve on the cfg3 stream: the e2e leg of
bench.py alone, for A/B runs of the host-path knobs (SPB_HOST_CHUNK_LOG2).  usage: python tools/e2e_probe.py [log2_points]"""
import sys
import time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sciport_rs_b200 import special  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = 1 << log2n
dev = torch.device("cuda", 0)
v, z = special.synth(3, n, n_total=1 << 28, device=dev)
vh = torch.empty(n, dtype=torch.float64).pin_memory()
zh = torch.empty(n, dtype=torch.complex128).pin_memory()
vh.copy_(v)
zh.copy_(z)
oi = torch.empty(n, dtype=torch.complex128).pin_memory()
ok = torch.empty(n, dtype=torch.complex128).pin_memory()
del v, z
vn, zn, oin, okn = vh.numpy(), zh.numpy(), oi.numpy(), ok.numpy()


def step(m=None):
    try:
        special.ive_kve(vn[:m], zn[:m], outs=[oin[:m], okn[:m]])
    except special.SpecialError:
        pass


step(1 << 23)
torch.cuda.synchronize()
best = None
for _ in range(3):
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print(f"e2e n=2^{log2n}: {best * 1e3:.1f} ms  {2 * n / best:.3e} evals/s  {56.0 * n / best / 1e9:.1f} GB/s over PCIe  "
      f"(SPB_HOST_CHUNK_LOG2={os.environ.get('SPB_HOST_CHUNK_LOG2', 'default 22')})")
