"""Warm per-launch device times of one steady-state frame (CUDA events around every engine step, serialised on one
stream: no overlap between branches, so the groups add up to more than the frame time of bench.py)."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import weights as W  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402

size = int(os.environ.get('PROFILE_SIZE', '512'))
weights, _ = bench.load_weights()
trk = bench.make_tracker(weights)
frames = [torch.from_numpy(f).cuda() for f in synthetic_video(bench.STEADY + 8, size, size, seed=1234)]
trk.init(frames[0].cpu().numpy())
t = 1
for _ in range(bench.STEADY + 2):
    trk.track(frames[t], device_result=True); t += 1
eng = trk.engine
reps = 3
groups = collections.OrderedDict()
for r in range(reps):
    eng.set_option('profile', 1)
    trk.track(frames[t], device_result=True); t += 1
    st = eng.profile_steps()
    eng.profile_fetch()
    eng.set_option('profile', 0)
    seen_prog = False
    if r == 0 and os.environ.get('STEP_DETAIL'):
        print('per step (us, kind 0 = tensor-core launch, layer tag):', ' '.join(f'{ms * 1e3:.1f}/{kind}/{layer}' for ms, kind, layer in st))
    for ms, kind, layer in st:
        if layer == 200 or layer == 201:
            name, seen_prog = 'iteration program (conv_prog_kernel)', True
        elif layer == 100:
            name = 'correlation GEMM'
        elif 0 <= layer < 16:
            name = 'fnet convs'
        elif 16 <= layer < 32:
            name = 'cnet convs'
        elif 0 <= layer < len(W.LAYER_NAMES):
            name = 'heads: ' + W.LAYER_NAMES[layer]
        else:
            name = 'other kernels after the loop (ou_pack, upsample)' if seen_prog else 'other kernels before / inside the loop (patches, instance norm, pair setup, pool, lookup)'
        g = groups.setdefault(name, [0.0, 0])
        g[0] += ms
        g[1] += 1
tot = sum(v[0] for v in groups.values()) / reps
print(f'{sum(v[1] for v in groups.values()) // reps} steps, serialised total {tot * 1e3:.0f} us (events add ~2-4 us per step)')
for k, (ms, n) in groups.items():
    print(f'{ms / reps * 1e3:9.1f} us  {n // reps:4d} launches  {k}')
