"""Host-side wall-clock breakdown of MFT.track() with host frames (where does e2e time go?)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mft_b200.MFT as M  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402

weights, _ = bench.load_weights()
trk = bench.make_tracker(weights)
frames = list(synthetic_video(bench.STEADY + 40, 512, 512, seed=1234))
trk.init(frames[0])
t = 1
for _ in range(bench.STEADY + 3):
    trk.track(frames[t]); t += 1

eng = trk.engine
acc = {}
def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0) + time.perf_counter() - t0; return r
    return w
eng.encode_frame = timed('encode_frame(call)', eng.encode_frame)
eng.refine = timed('refine(call)', eng.refine)
M.chain_select = timed('chain_select(call)', M.chain_select)
orig_sync = torch.cuda.current_stream().synchronize
N = 20
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    trk.track(frames[t]); t += 1
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f'e2e {tot / N * 1e3:.3f} ms/frame')
for k, v in acc.items():
    print(f'  {k}: {v / N * 1e3:.3f} ms')
# raw pieces
x = torch.empty((4, 512, 512), device='cuda')
t0 = time.perf_counter()
for _ in range(50):
    h = torch.empty((4, 512, 512), dtype=torch.float32, pin_memory=True)
print(f'pinned alloc {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms')
hs = [torch.empty((4, 512, 512), dtype=torch.float32, pin_memory=True) for _ in range(4)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50):
    hs[i % 4].copy_(x, non_blocking=True); torch.cuda.current_stream().synchronize()
print(f'D2H 4 MiB + sync {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms')
