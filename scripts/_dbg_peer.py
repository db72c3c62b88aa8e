import os, sys, time
os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda_optical_flow_2_b200 import Context, WARP_BILINEAR
from cuda_optical_flow_2_b200.dist import NativeStrips
from bench import synth_pairs_torch
W, H, L, win, world = 1920, 1080, 3, 9, 2
dev = torch.device("cuda", 0)
ctx = Context(0)
prev, nxt, pitch = synth_pairs_torch(1, W, H, dev, 7)
ranks = [NativeStrips(ctx, W, H, L, win, world, rk, dev, WARP_BILINEAR, 1.0, 16, transport="local") for rk in range(world)]
NativeStrips.connect_local(ranks)
streams = [torch.cuda.Stream(dev) for _ in range(world)]
torch.cuda.synchronize()
for it in range(3):
    for ns, st in zip(ranks, streams):
        y0, y1 = ns.own_rows(0)
        t0 = time.perf_counter()
        ns.run(prev[0, y0:y1], nxt[0, y0:y1], st.cuda_stream)
        print(f"iter {it} rank {ns.rank} run() host time {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"iter {it} sync {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
    for ns, st in zip(ranks, streams):
        try:
            ns.check(st.cuda_stream); print("check ok", ns.rank)
        except Exception as e:
            print("check failed", ns.rank, e)
