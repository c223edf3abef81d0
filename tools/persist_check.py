"""Persistent layer-program kernel vs one launch per layer: bit-exact outputs, device time of one refinement, the
scheduler run-ahead sweep and the role / per-layer timers of the program kernel (engine option prog_timing)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import _lib, engine as E  # noqa: E402
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


def timed():
    eng.refine(lefts, rights)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = eng.refine(lefts, rights)
    b.record()
    torch.cuda.synchronize()
    eng.check_device()
    return a.elapsed_time(b) / reps, out


res = {}
for mode, name in ((0, 'one launch per layer'), (1, 'one program per iteration'), (2, 'whole refinement in one launch')):
    eng.set_option('persist', mode)
    ms, out = timed()
    res[mode] = out.clone()
    print(f'persist={mode} ({name}): {ms:.3f} ms per {pairs}-pair refinement, bit-identical to per-layer launches: '
          f'{bool(torch.equal(res[0], out))}', flush=True)

eng.set_option('persist', 1)
for tk in (1, 2, 3, 4):
    eng.set_option('prog_tickets', tk)
    ms, out = timed()
    print(f'scheduler run-ahead {tk} ticket(s): {ms:.3f} ms  identical={bool(torch.equal(out, res[0]))}')

# role timers of the last launch (MMA warp of every CTA) and per-layer spans
eng.set_option('prog_timing', 1)
ptr, nb = ctypes.c_void_p(), ctypes.c_size_t()
_lib.check(_lib.lib().mftb200_debug_buffer(eng.ctx, b'prog_timing', ctypes.byref(ptr), ctypes.byref(nb)), eng.ctx)
rt = ctypes.cdll.LoadLibrary('libcudart.so')
torch.cuda.synchronize()
rt.cudaMemset(ptr, 0, 65536)
rt.cudaMemset(ctypes.c_void_p(ptr.value + 4096 * 8), 0xff, 16 * 16)          # per-layer first hand-out = +inf
for l in range(16):
    rt.cudaMemset(ctypes.c_void_p(ptr.value + (4096 + 2 * l + 1) * 8), 0, 8)
eng.refine(lefts, rights)
buf = eng.debug_buffer('prog_timing', torch.int64, (8192,)).cpu().numpy().astype(float)
t = buf[:4096].reshape(256, 16)
t = t[t[:, 4] > 0]
us = lambda c: c / 1.965e3                                                     # SM clock 1965 MHz
busy = t[:, 0] - t[:, 1]
print(f'{len(t)} CTAs, last launch: MMA warp span {us(t[:, 0].min()):.1f}..{us(t[:, 0].max()):.1f} us, busy (not waiting for a '
      f'ticket) {us(busy.min()):.1f} / {us(busy.mean()):.1f} / {us(busy.max()):.1f} us, {t[:, 4].mean():.1f} tiles, '
      f'{t[:, 5].mean():.0f} stages per CTA')
print(f'   MMA warp waits: ticket {us(t[:, 1].mean()):.1f} us, accumulator free {us(t[:, 2].mean()):.1f} us, operands '
      f'{us(t[:, 3].mean()):.1f} us ({t[:, 3].sum() / t[:, 5].sum():.0f} cycles per stage); issue + MMA '
      f'{us((t[:, 0] - t[:, 1] - t[:, 2] - t[:, 3]).mean()):.1f} us')
lt = buf[4096:4096 + 24].reshape(12, 2)
names = ['convc1', 'convf1', 'convc2', 'convf2', 'convm', 'zr1', 'q1', 'zr2', 'q2', 'fh1', 'fh2']
t0 = min(lt[l, 0] for l in range(11) if lt[l, 1] > 0)
print('per layer, us after the first ticket of the refinement: first tile handed out (iteration 0) .. last tile complete (last iteration)')
for l, nme in enumerate(names):
    print(f'   {nme:7s} {(lt[l, 0] - t0) / 1e3:8.1f} .. {(lt[l, 1] - t0) / 1e3:8.1f}')
