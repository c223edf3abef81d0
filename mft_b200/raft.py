"""RAFTWrapper with the reference's module surface (MFT/raft.py of serycjon/MFT), backed by the
native engine.  compute_flow(src, dst, mode='flow') -> flow (2,H,W), {'occlusion','sigma','debug'}.

Unlike the reference, the per-frame encoders and the pair refinement are separate engine calls
(encode_frame / refine) so that the tracker encodes every frame once and batches all of a
frame's delta pairs into one refinement (SURVEY.md TL;DR: 21 encoder passes -> 2 per frame)."""
import logging

import numpy as np
import torch
import torch.nn.functional as F

from . import weights as _weights
from .engine import Engine

logger = logging.getLogger(__name__)

MAX_PAIRS = 7          # pairs per batched refinement (the 7 delta chains of configs/MFT_cfg.py)
TRACKER_SLOTS = 34     # template + 32 past frames + current (MFT.cleanup_memory keeps max-delta frames)
N_SLOTS = TRACKER_SLOTS + 2   # + two scratch slots for stand-alone compute_flow calls


class RAFTWrapper:
    def __init__(self, config):
        """config: flow config with .model (checkpoint path, or a state dict), .flow_iters,
        .raft_params (must describe the shipped architecture: non-small, separate OU heads)."""
        self.C = config
        params = getattr(config, 'raft_params', None)
        if params:
            small = params.get('small', False) if isinstance(params, dict) else getattr(params, 'small', False)
            occl = params.get('occlusion_module', None) if isinstance(params, dict) else getattr(params, 'occlusion_module', None)
            if small or occl != 'separate_with_uncertainty':
                raise NotImplementedError('mft_b200 implements the shipped RAFT-OU architecture only '
                                          "(small=False, occlusion_module='separate_with_uncertainty')")
        model = config.model
        state = model if isinstance(model, dict) else _weights.load_checkpoint(str(model))
        self.iters = int(config.flow_iters) if config.flow_iters else 12
        self._state = state
        self.engine = Engine(state)       # the tracker's engine: its feature slots mirror MFT.memory
        self._flow_engine = None          # stand-alone compute_flow calls of another geometry get their own workspace
        self._claimed = None              # geometry a tracker holds live feature slots for
        self.model = self.engine          # the reference exposes the network as .model (raft.py:28); here: the engine handle
        self.last_flow_shape = None

    def ensure_geometry(self, H, W, claim=False):
        """Engine configured for H x W frames.  claim=True (MFT.init): the tracker's engine, (re)configured freely --
        init resets all tracker state anyway.  Otherwise (compute_flow): the tracker's engine only if it already has
        this geometry; a different size goes to a second engine, because reconfiguring would zero the feature slots
        a running track depends on (the reference's compute_flow is stateless)."""
        if claim or self._claimed is None or self._claimed == (H, W):
            self.engine.configure(H, W, max_pairs=MAX_PAIRS, n_slots=N_SLOTS, iters=self.iters)
            if claim:
                self._claimed = (H, W)
            return self.engine
        if self._flow_engine is None:
            self._flow_engine = Engine(self._state)
        self._flow_engine.configure(H, W, max_pairs=1, n_slots=2, iters=self.iters)
        return self._flow_engine

    def compute_flow(self, src_img, dst_img, mode='TC', vis=False, src_img_identifier=None,
                     numpy_out=False, init_flow=None, vis_debug=False):
        """src_img, dst_img: (H,W,3) uint8 BGR.  mode 'flow' or 'TC' (raft.py:30-94)."""
        H, W = src_img.shape[:2]
        flow_init = None
        if init_flow is not None:
            # MFT/raft.py:49-53: replicate-pad like the images (InputPadder), then downsample_flow_8 -> RAFT's flow_init.
            # (Host-side preparation of an optional argument: two torch ops on a (2,H,W) tensor.)
            f = torch.as_tensor(init_flow).to('cuda', torch.float32).reshape(1, 2, H, W)
            ph, pw = (8 - H % 8) % 8, (8 - W % 8) % 8
            if ph or pw:
                f = F.pad(f, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), mode='replicate')
            flow_init = downsample_flow_8(f).contiguous()
        eng = self.ensure_geometry(H, W)
        s0 = TRACKER_SLOTS if eng is self.engine else 0        # the two scratch slots behind the tracker's ring
        in_place = eng.encode_frame(src_img, s0)
        in_place = eng.encode_frame(dst_img, s0 + 1) or in_place
        out = eng.refine([s0], [s0 + 1], init_flow=flow_init)[0]
        eng.error_flag_async()
        if in_place:
            eng.wait_frame_copied()
        eng.error_flag_poll()
        flow, occlusions, sigma = out[0:2], out[2:3], out[3:4]
        extra = {'occlusion': occlusions, 'sigma': sigma, 'debug': None}
        if mode == 'flow':
            if numpy_out:
                flow = flow.cpu().numpy()
                extra['occlusion'] = occlusions.cpu().numpy()
                extra['sigma'] = sigma.cpu().numpy()
            return flow, extra
        if mode == 'TC':
            self.last_flow_shape = {'delta': 2, 'H': H, 'W': W}
            ys, xs = torch.meshgrid(torch.arange(H, device=flow.device), torch.arange(W, device=flow.device), indexing='ij')
            src = torch.stack([xs.reshape(-1), ys.reshape(-1)], 0)
            dst = src + flow.reshape(2, H * W)
            if numpy_out:
                src, dst = src.cpu().numpy(), dst.cpu().numpy()
                extra['occlusion'] = occlusions.reshape(-1).cpu().numpy()
                extra['sigma'] = sigma.reshape(-1).cpu().numpy()
            return src, dst, extra
        raise ValueError(f'unknown mode {mode}')


def downsample_flow_8(flow, mode='bilinear'):
    """(B, xy, H, W) -> (B, xy, H/8, W/8), values / 8 (raft.py:98-101)."""
    size = (flow.shape[2] // 8, flow.shape[3] // 8)
    return F.interpolate(flow, size=size, mode=mode, align_corners=True) / 8
