import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import mft_oracle as O
from mft_b200 import engine as E
from mft_b200.synth import synthetic_video
W = O.seeded_weights(0)
frames = list(synthetic_video(3, 128, 160, seed=3))
eng = E.Engine(W)
eng.configure(128, 160, max_pairs=2, n_slots=3, iters=2)
for i, f in enumerate(frames):
    eng.encode_frame(f, i)
for mode in (1, 2):
    eng.set_option('persist', mode)
    out = eng.refine([0, 1], [2, 2])
    eng.check_device()
    print('mode', mode, bool(torch.isfinite(out).all()))
r = out.cpu()
lefts = [torch.zeros(4, 128, 160, device='cuda'), out[1].clone()]
res, idx = E.chain_select(lefts, out, 0.02)
print('done', float(res.abs().mean()))
