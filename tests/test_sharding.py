"""Multi-GPU host logic on CPU: world_size-2 gloo run of the shard partition + the throughput
all-gather, and the shard-invariance property (global env ids key the RNG) checked with the oracle."""
import os
import subprocess
import sys

import numpy as np

import oracle
import parity
from xworld_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from xworld_b200.sharding import shard_range, gather_throughput
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1001, r, w)
total, worst, per_rank = gather_throughput((hi - lo) * 10, 1000 + 500 * r)
assert total == 1001 * 10 and worst == 1000 + 500 * (w - 1) and len(per_rank) == w, (total, worst, per_rank)
print("rank", r, "ok", lo, hi)
dist.destroy_process_group()
"""


def test_shard_ranges_partition():
    for n in (1, 7, 4096, 65536, 262144 + 3):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_gloo_world_size_2():
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()


def test_sharded_batch_equals_single_batch(synthetic_catalog):
    """Two shards of a 64-env batch == the unsharded batch, env for env (oracle; the GPU test repeats
    it through the C ABI)."""
    n = 64
    full = oracle.Oracle(parity.make_cfg("c2_nav3d_7x7_84"), synthetic_catalog, n, threads=2)
    full.reset()
    shards = []
    for r in range(2):
        lo, hi = shard_range(n, r, 2)
        o = oracle.Oracle(parity.make_cfg("c2_nav3d_7x7_84", env_id_offset=lo), synthetic_catalog, hi - lo, threads=2)
        o.reset()
        shards.append((lo, hi, o))
    for s in range(30):
        a = parity.actions_for(s, n, 4)
        r_full, o_full, _ = full.step(a)
        for lo, hi, o in shards:
            r, ov, _ = o.step(a[lo:hi])
            assert (r.view(np.uint32) == r_full[lo:hi].view(np.uint32)).all() and (ov == o_full[lo:hi]).all()
    for lo, hi, o in shards:
        assert (o.field("grid") == full.field("grid")[lo:hi]).all()


def test_bench_clock_sampler_picks_rows_of_the_timed_region():
    """bench.py's ClockSampler: rows stamped inside [begin, end + one period] are the ones reported; a region shorter
    than a period falls back to the rows next to it; no nvidia-smi (this container) is reported as such."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc(object):
        def terminate(self):
            pass

    class FakeThread(object):
        def join(self, timeout=None):
            pass

    def sampler(stamps, clocks, t0, t1):
        s = bench.ClockSampler(0)
        s.proc, s.t = FakeProc(), FakeThread()
        s.stamps = list(stamps)
        s.rows = [["0", str(c), "1965", "300", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"] for c in clocks]
        s.t0, s.t1 = t0, t1
        return s.stop()

    out = sampler([0.00, 0.02, 0.04, 0.06, 0.08, 0.50], [300, 1965, 1950, 1965, 1200, 210], 0.015, 0.065)
    assert out["window"] == "timed region" and out["samples"] == 4 and out["sm_mhz"] == 1957.5 and out["reasons"] == []
    out = sampler([0.00, 0.10, 0.20], [300, 1965, 400], 0.11, 0.12)
    assert out["window"].startswith("rows adjacent") and out["samples"] >= 2
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]
