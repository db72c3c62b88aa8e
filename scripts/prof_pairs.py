"""Run the device-resident path a few times on a small batch (target for ncu / quick timing).
usage: python scripts/prof_pairs.py [pairs] [reps] [w] [h] [levels] [win] [warp_mode]   (OFB_SOLVE=1: tolerance-mode solve)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_pairs_torch
from cuda_optical_flow_2_b200 import Context

a = [int(x) for x in sys.argv[1:]] + [None] * 7
pairs, reps, w, h, levels, win, mode = (a[0] or 8), (a[1] or 3), (a[2] or 1920), (a[3] or 1080), (a[4] or 3), (a[5] or 9), (2 if a[6] is None else a[6])
dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.solve = int(os.environ.get("OFB_SOLVE", "0"))
prev, nxt, pitch = synth_pairs_torch(pairs, w, h, dev, 1)
flows = [torch.empty((pairs, h >> k, w >> k, 2), dtype=torch.float32, device=dev) for k in range(levels)]
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ctx.flow_pairs_device(prev, nxt, w, levels, win, warp_mode=mode, flows=flows, stream=st)
torch.cuda.synchronize()
ctx.profile_enable(True)
t0 = time.perf_counter()
for _ in range(reps):
    ctx.flow_pairs_device(prev, nxt, w, levels, win, warp_mode=mode, flows=flows, stream=st)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"{pairs} pairs {w}x{h} L{levels} win{win} mode{mode} solve{ctx.solve}: {dt*1e3:.3f} ms/step, {w*h/1e6*pairs/dt:.0f} Mpx-pairs/s")
import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL)
for k in range(levels):
    ms, n = ctx.profile_read(k)
    npx = (w >> k) * (h >> k) * pairs
    print(f"  level {k}: {ms/n*1e3:.1f} us/launch, {npx/(ms/n*1e-3)/1e9:.1f} Gpx/s")
ms, n = ctx.profile_read(100)
print(f"  pyramid: {ms/max(n,1)*1e3:.1f} us")
