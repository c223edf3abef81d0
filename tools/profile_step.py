"""One steady-state tracked frame between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402

H, W = bench.parse_size(os.environ.get('PROFILE_SIZE', ''), 512)          # N or HxW
steps = int(os.environ.get('PROFILE_STEPS', '1'))
weights, _ = bench.load_weights()
trk = bench.make_tracker(weights)
frames = [torch.from_numpy(f).cuda() for f in synthetic_video(bench.STEADY + 4 + steps, H, W, seed=1234)]
trk.init(frames[0].cpu().numpy())
for kv in filter(None, os.environ.get('BENCH_ENGINE_OPTIONS', '').split(',')):      # same A/B knobs as bench.py
    k, v = kv.split('=')
    trk.engine.set_option(k.strip(), int(v))
t = 1
for _ in range(bench.STEADY + 2):
    trk.track(frames[t], device_result=True)
    t += 1
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    trk.track(frames[t], device_result=True)
    t += 1
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
trk.engine.check_device()
print('profiled', steps, 'step(s)')
