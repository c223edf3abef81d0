"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck are slow: two pairs, two iterations):
   compute-sanitizer --tool memcheck python tools/sanitizer_case.py
Two geometries: 128x160 (coarse width 20: gather lookup) and 128x192 (coarse width 24: TMA lookup with padded pyramid rows,
persistent correlation kernel with ragged last slice); per-iteration program and the all-iterations program."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mft_b200 import engine as E  # noqa: E402
from mft_b200 import weights as WT  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402

W = WT.random_init(0)
for H, Wd in ((128, 160), (128, 192)):
    frames = list(synthetic_video(3, H, Wd, seed=3))
    eng = E.Engine(W)
    eng.configure(H, Wd, max_pairs=2, n_slots=3, iters=2)
    for i, f in enumerate(frames):
        eng.encode_frame(f, i)
    for mode in (1, 2):
        eng.set_option('persist', mode)
        out = eng.refine([0, 1], [2, 2])
        eng.check_device()
        print(f'{H}x{Wd} persist {mode}: finite {bool(torch.isfinite(out).all())}')
    lefts = [torch.zeros(4, H, Wd, device='cuda'), out[1].clone()]
    res, idx = E.chain_select(lefts, out, 0.02)
    print(f'{H}x{Wd} done', float(res.abs().mean()))
