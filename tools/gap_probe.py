"""Where does the time between kernels go?  Times K pair calls (a) through special.bessel_jy, (b) through the raw
C ABI with preallocated outputs, and prints the per-call kernel sum from a profiled call."""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sciport_rs_b200 import special, _lib
from sciport_rs_b200._lib import lib, make_exec
n = 4096 * 4096
v, z = special.synth(2, n, n_total=n, device="cuda:0")
vt = torch.full((1,), 2.5, dtype=torch.float64, device="cuda")
outs = [torch.empty(n, dtype=torch.complex128, device="cuda") for _ in range(2)]
codes = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4)]
ex = make_exec(device_ptrs=True, stream=torch.cuda.current_stream().cuda_stream, devices=[0])
def raw():
    lib.spb_bessel_jy(1, vt.data_ptr(), 0, z.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), codes[0].data_ptr(),
                      codes[1].data_ptr(), codes[2].data_ptr(), codes[3].data_ptr(), n, ctypes.byref(ex))
def timeit(f, k=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): f()
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, 1e3 * t_host / k
print("raw C ABI      : %.3f ms/call device, %.3f ms/call host enqueue" % timeit(raw))
print("special.bessel_jy: %.3f ms/call device, %.3f ms/call host enqueue" % timeit(lambda: special.bessel_jy(vt, z)))
prof = special.profile_step(lambda: special.bessel_jy(vt, z, profile=True))
print("profiled kernel sum %.3f ms" % sum(r["ms"] for r in prof), [(r["name"], round(r["ms"], 3)) for r in prof])
