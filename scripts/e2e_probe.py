"""End-to-end host-pointer path with different pipeline shapes (OFB_E2E_LANES / OFB_E2E_SUB are read at context creation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cuda_optical_flow_2_b200 import Context, WARP_BILINEAR
Be = 32
ctx = Context(0)
hp, hn = bench.synth_pairs_numpy(2, bench.W, bench.H, 5)
pin = bench.Pinned()
hprev, hnext = pin.array((Be, bench.H, bench.W, 3), np.uint8), pin.array((Be, bench.H, bench.W, 3), np.uint8)
for i in range(Be):
    hprev[i], hnext[i] = hp[i % 2], hn[i % 2]
houts = [pin.array((Be, bench.H >> k, bench.W >> k, 2), np.float32) for k in range(3)]
for _ in range(2):
    ctx.flow_pairs_host(hprev, hnext, 3, 9, warp_mode=WARP_BILINEAR, out=houts)
t0 = time.perf_counter()
for _ in range(6):
    ctx.flow_pairs_host(hprev, hnext, 3, 9, warp_mode=WARP_BILINEAR, out=houts)
dt = (time.perf_counter() - t0) / 6
print(f"lanes {os.environ.get('OFB_E2E_LANES','3')} sub {os.environ.get('OFB_E2E_SUB','auto')}: {bench.W*bench.H/1e6*Be/dt:.0f} Mpx-pairs/s, D2H {sum(a.nbytes for a in houts)/dt/1e9:.1f} GB/s")
