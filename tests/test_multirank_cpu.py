"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path — contiguous slices with no shared element
(no collective on the data path), the max-over-ranks timing rule, and the reference arm under torchrun (rank 0
prints one JSON line, the other ranks exit 0 without work)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch = pytest.importorskip("torch")


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1000003
    lo, hi = bench.slice_bounds(n, world)[rank]
    # each rank marks its slice of a shared-size array; the union must cover [0, n) exactly once
    mark = torch.zeros(n, dtype=torch.int32)
    mark[lo:hi] = 1
    dist.all_reduce(mark)  # test-only reduction to verify the partition (the data path itself has none)
    slowest = bench.max_over_ranks(10.0 + 5.0 * rank)
    json.dump({"lo": lo, "hi": hi, "cover_min": int(mark.min()), "cover_max": int(mark.max()), "slowest": slowest},
              open(os.path.join(out_dir, f"r{rank}.json"), "w"))
    dist.destroy_process_group()


def test_slices_and_max_over_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    recs = [json.load(open(tmp_path / f"r{r}.json")) for r in range(world)]
    assert recs[0]["lo"] == 0 and recs[0]["hi"] == recs[1]["lo"] and recs[1]["hi"] == 1000003
    for r in recs:
        assert r["cover_min"] == 1 and r["cover_max"] == 1
        assert r["slowest"] == 15.0  # the slowest rank defines the step on every rank


def test_slice_bounds_properties():
    sys.path.insert(0, ROOT)
    import bench
    for n in (0, 1, 7, 4096 * 4096, 10**9 + 7):
        for g in (1, 2, 3, 4, 8):
            b = bench.slice_bounds(n, g)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_reference_arm_under_torchrun_two_ranks():
    """bench.py --impl reference launched like the driver does for N = 2: one JSON line from rank 0 only."""
    port = 31000 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--ref-sample", "65536"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["e2e"]["h2d_bytes_per_step"] == 0
