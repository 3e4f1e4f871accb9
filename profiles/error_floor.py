"""BASELINE config 5: error-floor sweep of the n=1024 sample code down to FER ~1e-7 and below on the BSC and the BEC,
--frame-error-count 100, through the sweep driver; one process per GPU under torchrun (frames of every round sharded over
the ranks, counters all-reduced), or a single process.
usage: [torchrun --nproc-per-node N] python profiles/error_floor.py [bsc|bec ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    from libldpc_b200 import dist as D
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
ctx = api.Context(H, "", device=local)
which = sys.argv[1:] or ["bsc", "bec"]
plan = {"bsc": ("BSC", "BP_MS", [0.17, 0.16, 0.15, 0.14], 4_000_000_000), "bec": ("BEC", "BP", [0.82, 0.80, 0.78, 0.76, 0.74], 40_000_000_000)}
for key in which:
    ch, dec, xs, cap = plan[key]
    if rank == 0:
        print(f"\n### {ch}, {dec}, -i 50, early termination, --frame-error-count 100, --max-frames {cap:.0e}, {world} GPU(s)\n")
        print("| eps | frames | frame errors | FER | BER | avg iters | wall s | frames/s |")
        print("|---|---|---|---|---|---|---|---|", flush=True)
    for x in xs:
        t0 = time.perf_counter()
        kw = dict(channel=ch, decoding=dec, iterations=50, early_term=True, max_frames=cap, fec=100, seed=2)
        r = D.simulate_distributed(ctx, [x, x + 1e-9, 1.0], **kw) if world > 1 else ctx.simulate([x, x + 1e-9, 1.0], **kw)
        dt = time.perf_counter() - t0
        if rank == 0:
            if len(r["frames"]):
                print(f"| {x:g} | {int(r['frames'][0])} | {int(r['fec'][0])} | {float(r['fer'][0]):.3e} | {float(r['ber'][0]):.3e} | {float(r['avg_iter'][0]):.2f} | "
                      f"{dt:.1f} | {int(r['frames'][0]) / dt:.3e} |", flush=True)
            else:
                print(f"| {x:g} | {cap} | 0 | < {1 / cap:.1e} | - | - | {dt:.1f} | {cap / dt:.3e} |", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
