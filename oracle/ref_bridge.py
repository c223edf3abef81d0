"""Bridge to the UNMODIFIED reference (serycjon/MFT) running on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where the read-only reference checkout exists (the build container:
/root/reference).  Used by ``oracle/make_golden.py`` to produce ``tests/golden/*.npz`` and by
``tests/test_oracle_vs_reference.py`` to validate ``oracle/mft_oracle.py`` directly.
Nothing under ``mft_b200/``, no ``-m gpu`` test, ``smoke()`` or ``bench.py`` imports this.

The reference hard-codes the device string 'cuda' in exactly two places
(MFT/MFT.py:20, MFT/raft.py:17,25,45,50).  The two subclasses below avoid those lines and
reuse everything else (MFT.track, chain_results, FlowOUTrackingResult, RAFT) as is
(recipe from SURVEY.md §8c).
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

def _find_root():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.environ.get('MFT_REFERENCE_ROOT', ''), '/root/reference', os.path.join(here, 'baseline', '_ref')):
        if p and os.path.isfile(os.path.join(p, 'MFT', 'MFT.py')):
            return p
    return os.environ.get('MFT_REFERENCE_ROOT', '/root/reference')


REF_ROOT = _find_root()
CKPT_REL = 'checkpoints/raft-things-sintel-kubric-splitted-occlusion-uncertainty-non-occluded-base-sintel.pth'
VIDEO_REL = 'demo_in/ugsJtsO9w1A-00.00.24.457-00.00.29.462_HD.mp4'


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'MFT', 'MFT.py'))


@contextlib.contextmanager
def _in_ref_root():
    old = os.getcwd()
    os.chdir(REF_ROOT)          # configs + checkpoint paths are CWD-relative (configs/MFT_cfg.py:14)
    try:
        yield
    finally:
        os.chdir(old)


def _import_reference():
    sys.dont_write_bytecode = True      # read-only mount
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import MFT.MFT as ref_mft          # noqa
    import MFT.results as ref_results  # noqa
    import MFT.config as ref_config    # noqa
    from MFT.RAFT.core.raft import RAFT
    from MFT.RAFT.core.utils.utils import InputPadder
    return ref_mft, ref_results, ref_config, RAFT, InputPadder


def build_reference_model(state_dict=None):
    """The reference RAFT module on CPU.  state_dict=None -> shipped checkpoint; otherwise a
    flat {name: tensor} dict without the 'module.' prefix (e.g. oracle.seeded_weights())."""
    _, _, _, RAFT, _ = _import_reference()

    class AttrDict(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)

    args = AttrDict(occlusion_module='separate_with_uncertainty', small=False, mixed_precision=False)
    model = RAFT(args)
    if state_dict is None:
        sd = torch.load(os.path.join(REF_ROOT, CKPT_REL), map_location='cpu', weights_only=True)
        sd = {k[len('module.'):]: v for k, v in sd.items()}
        model.load_state_dict(sd)
    else:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        # cnet registers norm3 twice (as .norm3 and .downsample.1); BN counters are unused
        missing = [m for m in missing if 'num_batches_tracked' not in m and '.downsample.1.' not in m]
        assert not missing and not unexpected, (missing, unexpected)
        for li in (2, 3):                        # alias -> same module, nothing to do
            blk = getattr(model.cnet, f'layer{li}')[0]
            assert blk.downsample[1] is blk.norm3
    model.requires_grad_(False)
    model.eval()
    return model


class CpuFlower:
    """RAFTWrapper.compute_flow(mode='flow') (MFT/raft.py:39-73) minus the .cuda() calls."""

    def __init__(self, model, iters=12):
        self.model, self.iters = model, iters
        _, _, _, _, self.InputPadder = _import_reference()

    @torch.no_grad()
    def compute_flow(self, src_img, dst_img, mode='flow', init_flow=None, **kw):
        assert mode == 'flow'
        H, W = src_img.shape[:2]
        im1 = torch.from_numpy(src_img[:, :, ::-1].copy()).permute(2, 0, 1)[None].float()
        im2 = torch.from_numpy(dst_img[:, :, ::-1].copy()).permute(2, 0, 1)[None].float()
        padder = self.InputPadder(im1.shape)
        im1, im2 = padder.pad(im1, im2)
        if init_flow is not None:                      # MFT/raft.py:49-53 with the reference's own helpers
            from MFT.raft import downsample_flow_8
            init_flow, = padder.pad(torch.as_tensor(init_flow, dtype=torch.float32)[None])
            init_flow = downsample_flow_8(init_flow)
        pred = self.model(im1, im2, iters=self.iters, test_mode=True, flow_init=init_flow)
        flow = padder.unpad(pred['flow'])[0]
        occ = torch.squeeze(padder.unpad(pred['occlusion'].softmax(dim=1)[:, 1:2]), dim=0)
        sigma = torch.sqrt(torch.exp(torch.squeeze(padder.unpad(pred['uncertainty']), dim=0)))
        return flow, {'occlusion': occ, 'sigma': sigma, 'debug': None, 'raw': pred}


def build_reference_tracker(model, deltas, occlusion_threshold=0.02, iters=12, flower=None):
    """The reference MFT tracker class, on CPU, around ``model`` (or around ``flower``, any object with the
    reference's compute_flow(mode='flow') surface, e.g. a recording wrapper of CpuFlower)."""
    ref_mft, _, ref_config, _, _ = _import_reference()

    class CpuMFT(ref_mft.MFT):
        def __init__(self, config, flower):
            self.C = config
            self.flower = flower
            self.device = 'cpu'

    C = ref_config.Config()
    C.deltas = list(deltas)
    C.occlusion_threshold = occlusion_threshold
    return CpuMFT(C, flower if flower is not None else CpuFlower(model, iters))


def demo_frames(n, size=None, start=0):
    """First ``n`` frames of the reference's demo video as uint8 BGR, optionally resized with
    cv2.INTER_AREA to (W,H)=size."""
    import cv2
    cap = cv2.VideoCapture(os.path.join(REF_ROOT, VIDEO_REL))
    frames = []
    i = 0
    while len(frames) < n:
        ok, f = cap.read()
        if not ok:
            break
        if i >= start:
            if size is not None:
                f = cv2.resize(f, size, interpolation=cv2.INTER_AREA)
            frames.append(np.ascontiguousarray(f))
        i += 1
    cap.release()
    return frames
