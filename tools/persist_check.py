"""Persistent layer-program kernel vs one launch per layer: bit-exact outputs + device time of one refinement."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import engine as E  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402

size = int(os.environ.get('CHECK_SIZE', '512'))
pairs = int(os.environ.get('CHECK_PAIRS', '7'))
reps = int(os.environ.get('CHECK_REPS', '5'))
weights, _ = bench.load_weights()
eng = E.Engine(weights)
eng.configure(size, size, max_pairs=pairs, n_slots=pairs + 2, iters=12)
frames = list(synthetic_video(pairs + 1, size, size, seed=5))
for i, f in enumerate(frames):
    eng.encode_frame(f, i)
lefts, rights = list(range(pairs)), [pairs] * pairs
res = {}
for mode in (0, 1, 0, 1):
    eng.set_option('persist', mode)
    out = eng.refine(lefts, rights)
    eng.check_device()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = eng.refine(lefts, rights)
    b.record()
    torch.cuda.synchronize()
    eng.check_device()
    ms = a.elapsed_time(b) / reps
    print(f'persist={mode}: {ms:.3f} ms per {pairs}-pair refinement, finite={bool(torch.isfinite(out).all())}', flush=True)
    if mode in res:
        print(f'   repeatable: {bool(torch.equal(res[mode], out))}')
    res[mode] = out.clone()
d = (res[0] - res[1]).abs()
print(f'persist vs layered: bit-identical={bool(torch.equal(res[0], res[1]))} max|diff|={d.max().item():.3e} '
      f'flow mean|diff|={d[:, :2].mean().item():.3e}')

eng.set_option('persist', 1)
for tk in (1, 2, 3, 4):
    eng.set_option('prog_tickets', tk)
    eng.refine(lefts, rights)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        o2 = eng.refine(lefts, rights)
    b.record()
    torch.cuda.synchronize()
    eng.check_device()
    print(f'prog_tickets={tk}: {a.elapsed_time(b) / reps:.3f} ms  identical={bool(torch.equal(o2, res[1]))}')
import ctypes
eng.set_option('prog_timing', 1)
torch.cuda.synchronize()
z = torch.zeros(8192, dtype=torch.int64, device='cuda')
from mft_b200 import _lib
ptr = ctypes.c_void_p(); nb = ctypes.c_size_t()
_lib.check(_lib.lib().mftb200_debug_buffer(eng.ctx, b'prog_timing', ctypes.byref(ptr), ctypes.byref(nb)), eng.ctx)
rt = ctypes.cdll.LoadLibrary('libcudart.so')
rt.cudaMemset(ptr, 0, 65536)
rt.cudaMemset(ctypes.c_void_p(ptr.value + 4096 * 8), 0xff, 16 * 16)      # per-layer first start = +inf
for l in range(16):
    rt.cudaMemset(ctypes.c_void_p(ptr.value + (4096 + 2 * l + 1) * 8), 0, 8)
out = eng.refine(lefts, rights)
t = eng.debug_buffer('prog_timing', torch.int64, (256, 16)).cpu().numpy().astype(float)
t = t[t[:, 4] > 0]
us = lambda c: c / 1.965e3
print(f'{len(t)} CTAs; MMA warp span min {us(t[:, 0].min()):.1f} / max {us(t[:, 0].max()):.1f} us; busy (not waiting for tickets) min {us((t[:, 0] - t[:, 1]).min()):.1f} / mean {us((t[:, 0] - t[:, 1]).mean()):.1f} / max {us((t[:, 0] - t[:, 1]).max()):.1f} us; tiles min {t[:, 4].min():.0f} / max {t[:, 4].max():.0f}')
print(f'{len(t)} CTAs; MMA warp: {us(t[:, 0].mean()):.1f} us in the launch, {t[:, 4].mean():.1f} tiles, {t[:, 5].mean():.0f} stages per CTA')
print(f'   MMA warp waits: ticket {us(t[:, 1].mean()):.1f} us, accumulator free {us(t[:, 2].mean()):.1f} us, '
      f'operands {us(t[:, 3].mean()):.1f} us ({t[:, 3].sum() / t[:, 5].sum():.0f} cycles per stage); '
      f'issue+rest {us((t[:, 0] - t[:, 1] - t[:, 2] - t[:, 3]).mean()):.1f} us')
print(f'   lookup tiles: {t[:, 9].mean():.1f} per CTA, {us(t[:, 8].sum() / max(t[:, 9].sum(), 1)):.1f} us each (epilogue warp 4)')
lt = eng.debug_buffer('prog_timing', torch.int64, (8192,)).cpu().numpy()[4096:4096 + 24].reshape(12, 2).astype(float)
names = ['convc1', 'convf1', 'convc2', 'convf2', 'convm', 'zr1', 'q1', 'zr2', 'q2', 'fh1', 'fh2']
t0 = min(lt[l, 0] for l in range(11) if lt[l, 1] > 0)
print('per layer, us after the first ticket of the refinement: first tile handed out (iteration 0) .. last tile complete (last iteration)')
for l, nme in enumerate(names):
    print(f'   {nme:7s} {(lt[l, 0] - t0) / 1e3:7.1f} .. {(lt[l, 1] - t0) / 1e3:7.1f}')
