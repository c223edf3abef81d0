"""Device time of the phases of one steady-state frame (events around the three engine calls of MFT.track)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import engine as E  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402

size = int(os.environ.get('PROFILE_SIZE', '512'))
weights, _ = bench.load_weights()
eng = E.Engine(weights)
eng.configure(size, size, max_pairs=7, n_slots=9, iters=12)
frames = [torch.from_numpy(f).cuda() for f in synthetic_video(9, size, size, seed=1234)]
for i, f in enumerate(frames):
    eng.encode_frame(f, i)
lefts, rights = list(range(7)), [8] * 7
out = eng.refine(lefts, rights)
prev = [torch.zeros(4, size, size, device='cuda') for _ in range(7)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, reps=10):
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


print(f'encode_frame (fnet || cnet)      {timed(lambda: eng.encode_frame(frames[8], 8)):8.1f} us')
print(f'raft_refine (7 pairs, 12 iters)  {timed(lambda: eng.refine(lefts, rights, out=out)):8.1f} us')
print(f'chain_select (7 chains)          {timed(lambda: E.chain_select(prev, out, 0.02, want_index=False)):8.1f} us')
eng.check_device()
