import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bench import synth_pairs_torch
from cuda_optical_flow_2_b200 import Context
dev = torch.device("cuda", 0)
ctx = Context(0)
pairs, w, h, levels, win = 4, 1920, 1080, 3, 9
prev, nxt, pitch = synth_pairs_torch(pairs, w, h, dev, 1)
flows = ctx.flow_pairs_device(prev, nxt, w, levels, win, warp_mode=2)
torch.cuda.synchronize()
f2 = flows[2].cpu().numpy(); f1 = flows[1].cpu().numpy()
# cumulative flow at level 1 = 2*up(f2) + f1
up = np.repeat(np.repeat(f2, 2, axis=1), 2, axis=2)[:, :f1.shape[1], :f1.shape[2]]
cum1 = 2 * up + f1
for name, c in (("level2 flow (cum for level 1)", f2), ("cum level1 (for level 0)", cum1)):
    d = 2 * c  # displacement in finer-level pixels
    fin = np.isfinite(d).all(-1)
    mag = np.abs(np.where(np.isfinite(d), d, 0)).max(-1)
    med = np.nanmedian(np.where(fin[..., None], d, np.nan).reshape(pairs, -1, 2), axis=1)
    dev_ = np.abs(np.where(np.isfinite(d), d, 0) - med[:, None, None, :]).max(-1)
    print(name, "nonfinite %.4f%%" % (100 * (~fin).mean()), "median", med[0],
          "| >8px from median: %.3f%%" % (100 * (dev_ > 8).mean()), ">32: %.3f%%" % (100 * (dev_ > 32).mean()),
          ">1000: %.3f%%" % (100 * (mag > 1000).mean()), ">32768: %.4f%%" % (100 * (mag > 32768).mean()))
