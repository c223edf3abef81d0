"""Drop-in proof under the reference's OWN configuration path.

The two config files below are the reference's configs/MFT_cfg.py:1-19 and
configs/flow/RAFTou_kubric_huber_split_nonoccl.py:1-33 with exactly the edit INTEGRATION.md §1 prescribes (the imports of
MFT / RAFTWrapper / Config point at mft_b200) -- the AttrDict raft_params, the CWD-relative nested load_config and the
CWD-relative checkpoint path are the reference's.  The test then replays demo.py:44-69 (load_config ->
config.tracker_class(config) -> init / track -> convert_to_point_tracking(meta.result, cuda_queries) -> result.cpu())
and demo.py:130-146 (warp_forward with a mask), re-assigns tracker.C like run_MFT_tapvid.py:151, and checks the tracked
fields against the vectors recorded from the unmodified reference tracker."""
import os

import numpy as np
import pytest
import torch

from conftest import golden, record_parity

pytestmark = pytest.mark.gpu

MFT_CFG = '''from mft_b200.MFT import MFT                      # was: from MFT.MFT import MFT
from pathlib import Path
from mft_b200.config import Config, load_config    # was: from MFT.config import Config, load_config
import numpy as np

import logging
logger = logging.getLogger(__name__)


def get_config():
    conf = Config()

    conf.tracker_class = MFT
    conf.flow_config = load_config('configs/flow/RAFTou_kubric_huber_split_nonoccl.py')
    conf.deltas = [np.inf, 1, 2, 4, 8, 16, 32]
    conf.occlusion_threshold = 0.02

    conf.name = Path(__file__).stem
    return conf
'''

FLOW_CFG = '''from pathlib import Path
from mft_b200.config import Config                 # was: from MFT.config import Config
from mft_b200.raft import RAFTWrapper              # was: from MFT.raft import RAFTWrapper


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super(AttrDict, self).__init__(*args, **kwargs)
        self.__dict__.update(kwargs)


def get_config():
    conf = Config()

    conf.of_class = RAFTWrapper
    conf_name = Path(__file__).stem

    raft_kwargs = {
        'occlusion_module': 'separate_with_uncertainty',
        'small': False,
        'mixed_precision': False,
    }
    conf.raft_params = AttrDict(**raft_kwargs)
    # original model location:
    conf.model = 'checkpoints/raft-things-sintel-kubric-splitted-occlusion-uncertainty-non-occluded-base-sintel.pth'

    conf.flow_iters = 12

    conf.flow_cache_dir = Path(f'flow_cache/{conf_name}/')
    conf.flow_cache_ext = '.flowouX16.pkl'
    conf.name = Path(__file__).stem

    return conf
'''


def _get_queries(frame_shape, spacing):          # demo.py:109-120
    H, W = frame_shape
    xs, ys = np.meshgrid(np.arange(0, W, spacing), np.arange(0, H, spacing))
    return torch.from_numpy(np.vstack((xs.flatten(), ys.flatten())).T).float().cuda()


def test_reference_config_path_and_demo_sequence(tmp_path, monkeypatch):
    from mft_b200 import weights as WT
    from mft_b200.config import load_config
    from mft_b200.point_tracking import convert_to_point_tracking
    ckpt = WT.find_checkpoint()
    if ckpt is None:
        pytest.skip('shipped checkpoint not available on this box')
    (tmp_path / 'configs' / 'flow').mkdir(parents=True)
    (tmp_path / 'checkpoints').mkdir()
    (tmp_path / 'configs' / 'MFT_cfg.py').write_text(MFT_CFG)
    (tmp_path / 'configs' / 'flow' / 'RAFTou_kubric_huber_split_nonoccl.py').write_text(FLOW_CFG)
    os.symlink(ckpt, tmp_path / 'checkpoints' / WT.CKPT_NAME)
    monkeypatch.chdir(tmp_path)                     # the reference runs from its checkout root: every path above is CWD-relative

    config = load_config('configs/MFT_cfg.py')      # demo.py:48
    assert config.name == 'MFT_cfg' and config.flow_config.name == 'RAFTou_kubric_huber_split_nonoccl'
    assert not config.timers_enabled and not config.cache_delta_infinity          # missing attributes read as falsy
    tracker = config.tracker_class(config)          # demo.py:50
    assert tracker.C is config and tracker.flower.model is not None

    g = golden('track_real_128.npz')                # 10 demo frames at 128x128, recorded with deltas [inf,1,2,4,8]
    frames = g['frames']
    # the eval runner swaps the config object between runs and expects the change to take effect (run_MFT_tapvid.py:151)
    other = load_config('configs/MFT_cfg.py')
    other.deltas = g['deltas'].tolist()
    tracker.C = other

    results, queries, initialized = [], None, False
    for frame in frames:                            # demo.py:58-69
        if not initialized:
            meta = tracker.init(frame)
            initialized = True
            queries = _get_queries(frame.shape[:2], 30)
        else:
            meta = tracker.track(frame)
        coords, occlusions = convert_to_point_tracking(meta.result, queries)
        result = meta.result
        result.cpu()
        results.append((result, coords, occlusions))
        meta.frame_i = len(results) - 1             # callers attach attributes to meta (run_MFT_tapvid.py:281-282)
        assert not hasattr(meta, 'vis')

    n_q = int(queries.shape[0])
    for i, (result, coords, occlusions) in enumerate(results):
        assert not result.flow.is_cuda and tuple(result.flow.shape) == (2, 128, 128)
        assert tuple(result.occlusion.shape) == (1, 128, 128) and tuple(result.sigma.shape) == (1, 128, 128)
        assert isinstance(coords, np.ndarray) and coords.shape == (n_q, 2) and occlusions.shape == (n_q,)
        assert coords.dtype == np.float32 and occlusions.dtype == np.float32
        if i == 0:
            assert np.abs(coords - queries.cpu().numpy()).max() == 0 and np.abs(occlusions).max() == 0
        if f'result_{i}' in g.files:
            want = g[f'result_{i}']
            got = result.packed().numpy()
            epe = np.sqrt(((got[:2] - want[:2]) ** 2).sum(0))
            record_parity(f'dropin_track128_frame{i}', dict(epe_median=np.median(epe), epe_mean=epe.mean()))
            assert np.median(epe) < 0.0035, (i, np.median(epe))          # measured <= 0.0012 px
            # the point tracks are bilinear samples of that field (MFT/point_tracking.py:6-27)
            q = queries.cpu().numpy().astype(int)
            assert np.abs(coords - (q + got[:2, q[:, 1], q[:, 0]].T)).max() < 1e-4
    # demo.py:130-146: propagate an RGBA edit with the template-visible mask
    result = results[-1][0]
    edit = np.zeros((128, 128, 4), np.uint8)
    edit[40:80, 30:90] = (10, 200, 30, 255)
    visible = torch.logical_and((result.occlusion[0] < 0.5).cpu(), torch.from_numpy(edit[:, :, 3] > 0))
    premult = edit[:, :, :3].astype(np.float32) * (edit[:, :, 3:4].astype(np.float32) / 255.0)
    color = result.warp_forward(premult, mask=visible)
    alpha = result.warp_forward(edit[:, :, 3:4], mask=visible)
    assert isinstance(color, np.ndarray) and color.shape == (128, 128, 3) and alpha.shape == (128, 128, 1)
    assert np.isfinite(color).all() and float(alpha.max()) > 200
    # caller-side mutation of the returned result must not reach the tracker's stored state (the reference returns a clone, MFT.py:145)
    stored = tracker.memory[tracker.current_frame_i]['result'].packed().clone()
    dev = tracker.track(frames[-1], device_result=True)
    kept = tracker.memory[tracker.current_frame_i]['result'].packed().clone()
    dev.result.occlusion[:] = 1
    dev.result.cpu()
    assert tracker.memory[tracker.current_frame_i]['result'].packed().is_cuda
    assert torch.equal(tracker.memory[tracker.current_frame_i]['result'].packed(), kept) and stored.is_cuda
    tracker.engine.check_device()


def test_compute_flow_of_another_size_does_not_disturb_a_running_track(seeded_weights):
    """RAFTWrapper.compute_flow is stateless in the reference (MFT/raft.py:30-73): a stand-alone call on images of another
    size must not reconfigure the workspace whose feature slots a running track depends on."""
    from mft_b200.config import Config
    from mft_b200.MFT import MFT
    from mft_b200.raft import RAFTWrapper
    from mft_b200.synth import synthetic_video
    frames = list(synthetic_video(6, 128, 160, seed=3))
    other = list(synthetic_video(2, 136, 200, seed=4))

    def make():
        fc = Config(); fc.of_class = RAFTWrapper; fc.model = seeded_weights; fc.flow_iters = 12
        C = Config(); C.flow_config = fc; C.deltas = [np.inf, 1, 2]; C.occlusion_threshold = 0.02
        return MFT(C)
    a, b = make(), make()
    a.init(frames[0]); b.init(frames[0])
    for t in range(1, 6):
        ra = a.track(frames[t]).result.packed()
        if t == 3:
            flow, extra = b.flower.compute_flow(other[0], other[1], mode='flow')
            assert tuple(flow.shape) == (2, 136, 200) and torch.isfinite(flow).all()
        rb = b.track(frames[t]).result.packed()
        # (two encodes of one frame agree to round-off, not bit for bit: instance-norm statistics use atomics)
        assert float((ra[:2] - rb[:2]).abs().max()) < 0.05 and float((ra[2:] - rb[2:]).abs().max()) < 0.05, t
