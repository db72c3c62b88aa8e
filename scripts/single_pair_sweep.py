"""One pair alone: per-level kernel time against the rows-per-CTA split (OFB_LK_ROWS, read by csrc/lk_win.cu at
every launch), to check the cost model's choice in the regime where all CTAs of a launch are resident at once.
usage: python scripts/single_pair_sweep.py [w] [h] [levels] [win]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_pairs_torch
from cuda_optical_flow_2_b200 import Context

a = [int(x) for x in sys.argv[1:]] + [None] * 4
w, h, levels, win = (a[0] or 1920), (a[1] or 1080), (a[2] or 3), (a[3] or 9)
dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.solve = 1
prev, nxt, pitch = synth_pairs_torch(1, w, h, dev, 1)
flows = [torch.empty((1, h >> k, w >> k, 2), dtype=torch.float32, device=dev) for k in range(levels)]
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(tag):
    ctx.profile_enable(False)
    for _ in range(5):
        ctx.flow_pairs_device(prev, nxt, w, levels, win, warp_mode=2, flows=flows, stream=st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        ctx.flow_pairs_device(prev, nxt, w, levels, win, warp_mode=2, flows=flows, stream=st)
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / 100
    ctx.profile_enable(True)
    for _ in range(50):
        ctx.flow_pairs_device(prev, nxt, w, levels, win, warp_mode=2, flows=flows, stream=st)
    torch.cuda.synchronize()
    per = []
    for k in range(levels):
        ms, n = ctx.profile_read(k)
        per.append(ms / max(n, 1) * 1e3)
    ms, n = ctx.profile_read(100)
    print(f"{tag:>10}: pair {whole*1e3:6.1f} us | levels " + " ".join(f"{x:6.1f}" for x in per) + f" | pyramid {ms/max(n,1)*1e3:5.1f}", flush=True)


os.environ.pop("OFB_LK_ROWS", None)
run("model")
for rows in [int(x) for x in os.environ.get("SWEEP", "5 8 11 13 16 19 21 24 27 32 37 45 53 69 85 101 135 270").split()]:
    os.environ["OFB_LK_ROWS"] = str(rows)
    run(f"rows {rows}")
os.environ.pop("OFB_LK_ROWS", None)
run("model")
