"""Warm per-launch device times of one steady-state frame (CUDA events around every engine step)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
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
acc = None
reps = 3
for r in range(reps):
    eng.set_option('profile', 1)
    trk.track(frames[t], device_result=True); t += 1
    st = eng.profile_steps()
    eng.profile_fetch()
    eng.set_option('profile', 0)
    acc = [a + b[0] for a, b in zip(acc, st)] if acc else [b[0] for b in st]
kinds = [k for _, k, _ in st]
ENC = ['patches'] + [f'fnet{i}' for i in range(100)]
tot = sum(acc) / reps
print(f'{len(acc)} steps, total {tot * 1e3:.1f} us (events add ~2-4 us per step)')
names_iter = ['lookup', 'convc1', 'convc2', 'convf1', 'convf2', 'convm', 'zr1', 'q1', 'zr2', 'q2', 'fh1', 'fh2']
n_enc = len(acc) - 3 - 12 * 12 - 6
for i, v in enumerate(acc):
    us = v / reps * 1e3
    if i < n_enc:
        tag = f'enc[{i}]'
    elif i < n_enc + 3:
        tag = ['pair_setup', 'corr_gemm', 'corr_pool'][i - n_enc]
    elif i < n_enc + 3 + 144:
        j = i - n_enc - 3
        tag = f'it{j // 12}:{names_iter[j % 12]}'
        if j // 12 not in (0, 5, 11):
            continue
    else:
        tag = ['mask1', 'mask2', 'ou_pack', 'ou1', 'ou2', 'upsample'][i - n_enc - 147]
    print(f'{i:4d} {us:8.1f} us  kind={kinds[i]}  {tag}')
enc = sum(acc[:n_enc]) / reps * 1e3
it = sum(acc[n_enc + 3:n_enc + 147]) / reps * 1e3
print(f'encoders {enc:.0f} us, pre {sum(acc[n_enc:n_enc+3]) / reps * 1e3:.0f} us, 12 iterations {it:.0f} us, final {sum(acc[n_enc+147:]) / reps * 1e3:.0f} us')
