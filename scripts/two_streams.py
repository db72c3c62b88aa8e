"""Does splitting a device-resident batch over two streams (two contexts) fill the ragged tails of the level launches?
usage: python scripts/two_streams.py [pairs] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_pairs_torch
from cuda_optical_flow_2_b200 import Context

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w, h, levels, win = 1920, 1080, 3, 9
dev = torch.device("cuda", 0)
prev, nxt, pitch = synth_pairs_torch(pairs, w, h, dev, 1)
for nsplit in (1, 2, 3, 4):
    ctxs = [Context(0) for _ in range(nsplit)]
    for c in ctxs:
        c.solve = 1
    streams = [torch.cuda.Stream(dev) for _ in range(nsplit)]
    per = pairs // nsplit
    flows = [[torch.empty((per, h >> k, w >> k, 2), dtype=torch.float32, device=dev) for k in range(levels)] for _ in range(nsplit)]
    def step():
        for i, (c, s) in enumerate(zip(ctxs, streams)):
            c.flow_pairs_device(prev[i * per:(i + 1) * per], nxt[i * per:(i + 1) * per], w, levels, win, flows=flows[i], stream=s.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{nsplit} stream(s) x {per} pairs: {dt*1e3:.3f} ms/step, {w*h/1e6*per*nsplit/dt:.0f} Mpx-pairs/s")
    del ctxs, flows
