"""tests/golden/raft_init_flow.npz: RAFTWrapper.compute_flow(..., init_flow=...) of the UNMODIFIED reference on CPU
(MFT/raft.py:49-53: replicate pad like the images, downsample_flow_8, RAFT.forward's flow_init, core/raft.py:153-154).
TEST INFRASTRUCTURE.  Run in the build container:   python -m oracle.make_golden_init

Two cases: shipped checkpoint, demo frames 0 -> 8 at 128x128, initialised with the reference's own flow 0 -> 1 times 6
plus a constant; seeded weights at 131x140 (odd replicate pad) with a smooth synthetic initial flow.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mft_oracle as O      # noqa: E402
from oracle import ref_bridge as R      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def main():
    import cv2
    warnings.filterwarnings('ignore')
    assert R.available(), 'reference checkout not found'
    torch.set_num_threads(4)
    g = {}
    fr = R.demo_frames(10, size=(128, 128))
    real = R.CpuFlower(R.build_reference_model())
    f01, _ = real.compute_flow(fr[0], fr[1])
    init = (_np(f01) * 6.0 + np.array([0.75, -1.25], np.float32)[:, None, None]).astype(np.float32)
    flow, ex = real.compute_flow(fr[0], fr[8], init_flow=torch.from_numpy(init))
    g.update(real_frames=np.stack([fr[0], fr[8]]), real_init=init, real_flow=_np(flow), real_occ=_np(ex['occlusion']),
             real_sigma=_np(ex['sigma']), real_coords=_np(ex['raw']['coords'][0]))
    seeded = R.CpuFlower(R.build_reference_model(O.seeded_weights(0)))
    fp = [cv2.resize(f, (140, 131), interpolation=cv2.INTER_AREA) for f in R.demo_frames(3, size=(256, 256))]
    yy, xx = np.mgrid[0:131, 0:140].astype(np.float32)
    init = np.stack([3.0 * np.sin(yy / 17.0) + 0.02 * xx, 2.0 * np.cos(xx / 23.0) - 0.015 * yy]).astype(np.float32)
    flow, ex = seeded.compute_flow(fp[0], fp[2], init_flow=torch.from_numpy(init))
    g.update(pad_frames=np.stack([fp[0], fp[2]]), pad_init=init, pad_flow=_np(flow), pad_occ=_np(ex['occlusion']),
             pad_sigma=_np(ex['sigma']))
    path = os.path.join(OUT, 'raft_init_flow.npz')
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path))


if __name__ == '__main__':
    main()
