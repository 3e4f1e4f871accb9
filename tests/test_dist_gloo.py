"""Host-side multi-rank logic on CPU: two gloo ranks shard every round of the sweep, all-reduce the
counters and must reproduce the single-process totals, results and results file.  The frame
processor is injected through ldpc_b200_simulate_ex's round callback and is the ORACLE (test
infrastructure) here because this container has no GPU; on the GPU box the same driver calls the CUDA path."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import H_FILE, ROOT

WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from libldpc_b200 import api, dist as D
from oracle import oracle as O

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
ctx = api.Context({h!r}, "", device=-1)
oc = O.Code({h!r})
calls = []
def round_fn(point, x, frame0, n, counters, user):
    r = oc.sim_point("BSC", x, seed=7, point=point, frame0=frame0, nframes=n, decoding="BP_MS", iterations=15, threads=2)
    counters[0] += r["fec"]; counters[1] += r["bec"]; counters[2] += r["frames"]; counters[3] += r["iters"]
    calls.append((int(point), int(frame0), int(n)))
    return 0
res = ctx.simulate([0.18, 0.245, 0.02], channel="BSC", decoding="BP_MS", iterations=15, seed=7, max_frames=3000, fec=40,
                   result_file={out!r} + str(world), rank=rank, world=world,
                   allreduce=D.make_allreduce() if world > 1 else None, round_fn=round_fn)
print("RESULT" + json.dumps(dict(rank=rank, calls=calls, res={{k: [float(v) for v in a] for k, a in res.items()}})))
"""


def _run(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, h=H_FILE, out=str(tmp_path / "res_w")))
    procs = []
    import socket
    with socket.socket() as sk:          # a free port per run (a fixed one can linger in TIME_WAIT between the two runs)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    for r in range(world):
        env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=600)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads([l for l in o.splitlines() if l.startswith("RESULT")][0][6:]))
    return sorted(outs, key=lambda d: d["rank"])


def test_two_ranks_reproduce_single_rank(tmp_path):
    one = _run(1, tmp_path)[0]
    two = _run(2, tmp_path)
    # identical per-point results on every rank and equal to the single-process run
    for k in ("fer", "ber", "avg_iter", "fec", "frames"):
        assert two[0]["res"][k] == two[1]["res"][k] == one["res"][k], k
    assert len(one["res"]["fer"]) >= 3
    # every round's frame range is split contiguously, without gaps or overlap
    c1 = one["calls"]
    ca, cb = two[0]["calls"], two[1]["calls"]
    assert len(c1) == len(ca) == len(cb)
    for (p, f0, n), (pa, fa, na), (pb, fb, nb) in zip(c1, ca, cb):
        assert p == pa == pb and fa == f0 and fb == fa + na and na + nb == n
    # only rank 0 writes the results file; same layout (execution order = eps descending for BSC)
    f1 = (tmp_path / "res_w1").read_text().split("\n")
    f2 = (tmp_path / "res_w2").read_text().split("\n")
    assert f1[0] == f2[0] == "snr fer ber frames avg_iter frame_time"
    strip = lambda lines: [" ".join(l.split()[:5]) for l in lines]   # all but the wall-clock column
    assert strip(f1) == strip(f2)
    xs = [float(l.split()[0]) for l in f1[1:] if l]
    assert xs == sorted(xs, reverse=True)


def test_shard_range_partition():
    from libldpc_b200.dist import shard_range
    for n in (0, 1, 7, 8192, 10**9 + 7):
        for world in (1, 2, 3, 8):
            edges = [shard_range(100, n, r, world) for r in range(world)]
            assert edges[0][0] == 100 and edges[-1][1] == 100 + n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
