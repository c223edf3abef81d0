"""Full-size golden vectors from the UNMODIFIED reference on CPU (the sizes bench.py measures).  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_big [demo|synth|big|all]

  tests/golden/track_demo_512.npz   BASELINE config 2: demo video resized to 512x512 (cv2.INTER_AREA), deltas
                                    [inf,1,2,4,8,16,32], 12 iterations, frames 1..40 through the reference tracker
                                    (MFT/MFT.py:55-154): per-frame statistics of every frame, the full field (stride 2)
                                    and the best-chain index map of frames 1, 8, 33, 40
  tests/golden/track_synth_512.npz  the same for bench.py's synthetic video (mft_b200.synth, seed 1234), frames 1..34:
                                    frames 33/34 are steady-state frames (7 live chains) of the benchmarked workload
  tests/golden/raft_1024_32it.npz   BASELINE config 4's pair shape: one 1024x1024 pair (synthetic frames 0 -> 8), 32 iterations,
                                    through the reference's compute_flow (MFT/raft.py:39-73, core/raft.py:97-259)
  The golden files hold the CRC32 of every input frame.  The frames are regenerated at test time (synthetic video: the
  seeded generator; demo video: decoded from the copy in oracle/_ref/, data like the checkpoint) and must reproduce the
  CRCs -- they did on the B200 boxes (same image, same cv2).  GOLDEN_SAVE_FRAMES=1 also writes oracle/_ref/frames_*.npy
  (git-ignored, travels to the GPU box) as a fallback for boxes where they would not.

The best-chain index is not returned by the reference (MFT.py:123-124 keeps it local).  It is recomputed here from the
reference's OWN per-delta flows (recorded at its flower) and its OWN stored left results with the reference's chain_results and
the selection lines of MFT.py:114-124, and the field it selects must equal the reference's result bit for bit.  The
oracle's chain / select run on the same operands beside it: index agreement and field difference are stored
(``oracle_vs_reference_<i>``), which pins the oracle's chain+select at full size.
"""
from __future__ import annotations

import os
import sys
import time
import warnings
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mft_oracle as O      # noqa: E402
from oracle import ref_bridge as R      # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
REFDATA = os.path.join(ROOT, 'oracle', '_ref')
DELTAS = [np.inf, 1, 2, 4, 8, 16, 32]
QS = (0.5, 0.9, 0.99)


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def field_stats(full):
    """Per-frame summary of a (4,H,W) field: channel means, quantiles of |flow| and sigma, occluded fraction."""
    mag = np.sqrt(full[0].astype(np.float64) ** 2 + full[1].astype(np.float64) ** 2)
    return np.concatenate([full.reshape(4, -1).astype(np.float64).mean(1), np.quantile(mag, QS),
                           np.quantile(full[3].astype(np.float64), QS), [(full[2] > 0.5).mean()]])


class RecordingFlower:
    def __init__(self, inner):
        self.inner, self.calls = inner, []

    def compute_flow(self, *a, **kw):
        flow, ex = self.inner.compute_flow(*a, **kw)
        self.calls.append((_np(flow), _np(ex['occlusion']), _np(ex['sigma'])))
        return flow, ex


def track_run(model, frames, keep, name):
    R._import_reference()
    from MFT.MFT import chain_results
    from MFT.results import FlowOUTrackingResult as FR
    rec = RecordingFlower(R.CpuFlower(model, 12))
    trk = R.build_reference_tracker(model, DELTAS, flower=rec)
    trk.init(frames[0])
    g = {'deltas': np.array(DELTAS), 'frame_crc': np.array([crc(f) for f in frames], np.uint32), 'quantiles': np.array(QS),
         'keep': np.array(keep)}
    stats = []
    for i in range(1, len(frames)):
        t0 = time.time()
        rec.calls.clear()
        live = O.live_chains(DELTAS, i, 0, 1)
        lefts = [tuple(_np(x) for x in (trk.memory[l]['result'].flow, trk.memory[l]['result'].occlusion,
                                        trk.memory[l]['result'].sigma)) for _, l in live] if i in keep else None
        r = trk.track(frames[i]).result
        full = np.concatenate([_np(r.flow), _np(r.occlusion), _np(r.sigma)], 0)
        stats.append(field_stats(full))
        if i in keep:
            assert len(rec.calls) == len(live)
            # the reference's own index: its chain_results on its own operands, then the selection lines MFT.py:114-124
            cref = [chain_results(FR(*(torch.from_numpy(a) for a in l)), FR(*(torch.from_numpy(a) for a in q)))
                    for l, q in zip(lefts, rec.calls)]
            scores = -torch.stack([c.sigma for c in cref], 0)
            scores[torch.stack([c.occlusion for c in cref], 0) > 0.02] = -float('inf')
            ridx = scores.max(dim=0, keepdim=True).indices[0, 0].numpy().astype(np.uint8)
            sel = np.stack([np.concatenate([_np(c.flow), _np(c.occlusion), _np(c.sigma)], 0) for c in cref])   # (K,4,H,W)
            picked = np.take_along_axis(sel, ridx[None, None].astype(np.int64).repeat(4, 1), 0)[0]
            same = (picked == full) | np.isnan(full)
            same[2] |= full[2] == 1.0                       # invalid-mask overwrite (results.py:258-264)
            assert same.all(), (name, i, int((~same).sum()))
            # the oracle's chain + select on the same operands: ATen's CPU grid_sample rounds its interpolation sum
            # differently from the oracle's defined order (1 ulp on a few % of the pixels), so near-ties may flip
            cands = [O.chain(l, q) for l, q in zip(lefts, rec.calls)]
            wf, wo, ws, oidx = O.select(cands, 0.02)
            mine = np.concatenate([wf, wo, ws], 0)
            agree = float((oidx == ridx).mean())
            ok = oidx == ridx
            maxdiff = float(np.nanmax(np.abs(mine - full)[:, ok])) if ok.any() else 0.0
            assert agree > 0.999 and maxdiff < 1e-3, (name, i, agree, maxdiff)
            g[f'oracle_vs_reference_{i}'] = np.array([agree, maxdiff])
            g[f'result_{i}'] = np.ascontiguousarray(full[:, ::2, ::2])
            g[f'index_{i}'] = ridx
            g[f'live_{i}'] = np.array([l for _, l in live])
            print(f'{name}: frame {i}: oracle index agreement {agree:.6f}, max field diff where equal {maxdiff:.2e}', flush=True)
        print(f'{name}: frame {i} ({len(live)} chains) {time.time() - t0:.1f}s', flush=True)
    g['stats'] = np.stack(stats)
    np.savez_compressed(os.path.join(OUT, f'track_{name}.npz'), **g)
    if os.environ.get('GOLDEN_SAVE_FRAMES'):          # only needed if the frames cannot be regenerated bit for bit on the GPU box
        np.save(os.path.join(REFDATA, f'frames_{name}.npy'), np.stack(frames))


def main():
    warnings.filterwarnings('ignore')
    assert R.available(), 'reference checkout not found'
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    torch.set_num_threads(int(os.environ.get('GOLDEN_THREADS', '8')))
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(REFDATA, exist_ok=True)
    real = R.build_reference_model()
    from mft_b200.synth import synthetic_video
    if what in ('big', 'all'):
        fr = list(synthetic_video(9, 1024, 1024, seed=1234))
        pair = [fr[0], fr[8]]
        t0 = time.time()
        flow, ex = R.CpuFlower(real, 32).compute_flow(pair[0], pair[1])
        full = np.concatenate([_np(flow), _np(ex['occlusion']), _np(ex['sigma'])], 0)
        np.savez_compressed(os.path.join(OUT, 'raft_1024_32it.npz'), result=np.ascontiguousarray(full[:, ::2, ::2]),
                            coords=_np(ex['raw']['coords'][0]), stats=field_stats(full), quantiles=np.array(QS),
                            frame_crc=np.array([crc(f) for f in pair], np.uint32), iters=np.array(32))
        if os.environ.get('GOLDEN_SAVE_FRAMES'):
            np.save(os.path.join(REFDATA, 'frames_1024.npy'), np.stack(pair))
        print(f'1024x1024 / 32 iterations: {time.time() - t0:.1f}s', flush=True)
    if what in ('synth', 'all'):
        track_run(real, list(synthetic_video(35, 512, 512, seed=1234)), (1, 8, 33, 34), 'synth_512')
    if what in ('demo', 'all'):
        track_run(real, R.demo_frames(41, size=(512, 512)), (1, 8, 33, 40), 'demo_512')
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
