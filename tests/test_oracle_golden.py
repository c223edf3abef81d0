"""Pin oracle/mft_oracle.py against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py, run on CPU in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import mft_oracle as O

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))

# fp32 CPU round-off of two independent implementations of the same 12-iteration recurrence
# (thread count alone moves the reference by 6e-5 px).
TOL_FLOW, TOL_OCC, TOL_SIGMA = 1e-3, 3e-4, 1e-3


def _check_pair(g, a, b, W, frames):
    flow, occ, sigma = O.compute_flow(W, frames[a], frames[b])
    assert np.abs(flow.numpy() - g[f'flow_{a}_{b}']).max() < TOL_FLOW
    assert np.abs(occ.numpy() - g[f'occ_{a}_{b}']).max() < TOL_OCC
    ref_s = g[f'sigma_{a}_{b}']
    assert (np.abs(sigma.numpy() - ref_s) / (1 + ref_s)).max() < TOL_SIGMA


def test_config1_real_256(real_weights):
    """BASELINE.json configs[0]: demo frames 0,1 at 256x256, delta {1}, 12 iters."""
    g = golden('raft_real_256.npz')
    fr = g['frames']
    assert g['flow_0_1'].shape == (2, 256, 256) and np.isfinite(g['flow_0_1']).all()
    _check_pair(g, 0, 1, real_weights, {0: fr[0], 1: fr[1]})


@pytest.mark.parametrize('tag', ['seeded', 'real'])
def test_raft_128(tag, request):
    W = request.getfixturevalue(f'{tag}_weights')
    g = golden(f'raft_{tag}_128.npz')
    frames = dict(zip(g['frame_ids'].tolist(), g['frames']))
    x = 2 * (O.bgr_to_input(frames[0]) / 255.0) - 1.0
    assert np.abs(O.basic_encoder(x, W, 'fnet')[0].numpy() - g['fnet_0']).max() < 2e-4
    assert np.abs(O.basic_encoder(x, W, 'cnet')[0].numpy() - g['cnet_0']).max() < 2e-4
    _check_pair(g, 0, 1, W, frames)
    _check_pair(g, 0, 8, W, frames)


def test_raft_padded_size_seeded(seeded_weights):
    """131x140: InputPadder's replicate pad (2|3 rows, 2|2 columns) and the unpad (core/utils/utils.py:9-24, MFT/raft.py:47-58)."""
    g = golden('raft_seeded_pad.npz')
    assert tuple(O.pad_amounts(131, 140)) == (2, 2, 2, 3)                 # (left, right, top, bottom)
    assert g['flow_0_2'].shape == (2, 131, 140)
    _check_pair(g, 0, 2, seeded_weights, {0: g['frames'][0], 2: g['frames'][1]})


def test_init_flow_vs_reference(real_weights, seeded_weights):
    """compute_flow(init_flow=...) (MFT/raft.py:49-53, core/raft.py:153-154): replicate pad, downsample_flow_8, flow_init."""
    g = golden('raft_init_flow.npz')
    for tag, W in (('real', real_weights), ('pad', seeded_weights)):
        fr = g[f'{tag}_frames']
        flow, occ, sigma = O.compute_flow(W, fr[0], fr[1], init_flow=g[f'{tag}_init'])
        assert np.abs(flow.numpy() - g[f'{tag}_flow']).max() < TOL_FLOW, tag
        assert np.abs(occ.numpy() - g[f'{tag}_occ']).max() < TOL_OCC, tag
        assert (np.abs(sigma.numpy() - g[f'{tag}_sigma']) / (1 + g[f'{tag}_sigma'])).max() < TOL_SIGMA, tag
    # the initialisation matters: without it the result differs visibly
    fr = g['real_frames']
    plain, _, _ = O.compute_flow(real_weights, fr[0], fr[1])
    assert np.abs(plain.numpy() - g['real_flow']).max() > 0.01


def test_chain_select_vs_reference_tracker():
    """Chaining + selection + invalid mask, against MFT.track itself fed with recorded flows."""
    g = golden('chain_select.npz')
    thr = float(g['thr'])
    for c in range(int(g['ncase'])):
        left, right, ref = g[f'c{c}_left'], g[f'c{c}_right'], g[f'c{c}_out']
        K = left.shape[0]
        cands = [O.chain((left[k, :2], left[k, 2:3], left[k, 3:4]),
                         (right[k, :2], right[k, 2:3], right[k, 3:4])) for k in range(K)]
        flow, occ, sig, idx = O.select(cands, thr)
        out = np.concatenate([flow, occ, sig], 0)
        # near-ties may legitimately flip between two fp32 implementations: demand that
        # (almost) every pixel agrees tightly and none of the engineered regions disagree.
        with np.errstate(invalid='ignore'):
            d = np.abs(out - ref)
            d[(out == ref)] = 0                      # inf == inf
        bad = (d > 1e-4).any(0)
        assert bad.mean() < 0.002, (c, bad.mean())
        assert not bad[:14].any(), c                  # rows 0..13 hold the edge-case regions
        assert (idx[:4] == 0).all()                   # all occluded -> first candidate
        assert (idx[5:7] == 0).all()                  # exact ties -> lowest index


def test_select_semantics():
    """torch.max semantics restated: first max wins, NaN wins, all -inf -> 0."""
    H, W = 2, 3
    z2, z1 = np.zeros((2, H, W), np.float32), np.zeros((1, H, W), np.float32)
    def cand(s, o=0.0):
        return (z2.copy(), z1 + np.float32(o), z1 + np.float32(s))
    assert (O.select([cand(1.0), cand(0.5), cand(0.5)], 0.02)[3] == 1).all()
    assert (O.select([cand(1.0, 0.5), cand(0.5, 0.5)], 0.02)[3] == 0).all()
    assert (O.select([cand(1.0), cand(np.nan), cand(0.1)], 0.02)[3] == 1).all()
    assert (O.select([cand(1.0), cand(0.1, 0.02)], 0.02)[3] == 1).all()      # strict '>' threshold
    assert (O.select([cand(1.0), cand(0.1, 0.020001)], 0.02)[3] == 0).all()


def test_chain_identity_is_inexact():
    """identity o r == r only to ~2e-5: the normalise/unnormalise round trip."""
    rng = np.random.default_rng(3)
    H, W = 33, 47
    r = ((rng.standard_normal((2, H, W)) * 4).astype(np.float32),
         rng.uniform(0, 1, (1, H, W)).astype(np.float32),
         rng.uniform(0, 3, (1, H, W)).astype(np.float32))
    ident = (np.zeros((2, H, W), np.float32), np.zeros((1, H, W), np.float32), np.zeros((1, H, W), np.float32))
    f, o, s = O.chain(ident, r)
    assert np.abs(f - r[0]).max() < 1e-4 and np.abs(o - r[1]).max() < 1e-4 and np.abs(s - r[2]).max() < 1e-4


def test_live_chain_bookkeeping():
    """RAFT calls per frame 1,2,3,3,4,... and 7 once t > 32."""
    deltas = [np.inf, 1, 2, 4, 8, 16, 32]
    n = [len(O.live_chains(deltas, t, 0, 1)) for t in range(1, 40)]
    assert n[:5] == [1, 2, 3, 3, 4] and n[32:] == [7] * 7
    assert O.live_chains(deltas, 5, 0, 1) == [(np.inf, 0), (1, 4), (2, 3), (4, 1)]
    assert O.live_chains(deltas, 95, 100, -1) == [(np.inf, 100), (1, 96), (2, 97), (4, 99)]


def test_lookup_channel_order_and_pyramid():
    """corr[n, v, u] = <f1[:, n], f2[:, v, u]>/sqrt(C); level l = avg-pooled target; lookup
    channel c <-> (dx = c//9 - 4, dy = c%9 - 4) (SURVEY §8c known-answer facts)."""
    rng = np.random.default_rng(5)
    h, w, C = 16, 24, 8
    f1 = torch.from_numpy(rng.standard_normal((1, C, h, w)).astype(np.float32))
    f2 = torch.from_numpy(rng.standard_normal((1, C, h, w)).astype(np.float32))
    pyr = O.corr_pyramid(f1, f2)
    assert [tuple(p.shape) for p in pyr] == [(h * w, h, w), (h * w, h // 2, w // 2), (h * w, h // 4, w // 4), (h * w, h // 8, w // 8)]
    n = 5 * w + 7
    want = (f1[0, :, 5, 7] * f2[0, :, 3, 11]).sum() / np.sqrt(C)
    assert abs(pyr[0][n, 3, 11] - want) < 1e-5
    pooled = torch.nn.functional.avg_pool2d(f2, 2)
    want1 = (f1[0, :, 5, 7] * pooled[0, :, 1, 4]).sum() / np.sqrt(C)
    assert abs(pyr[1][n, 1, 4] - want1) < 1e-5
    coords = O.coords_grid(h, w)
    out = O.corr_lookup(pyr, coords)
    c = (2 + 4) * 9 + (-1 + 4)          # dx=+2, dy=-1 at level 0
    assert abs(out[c, 5, 7] - pyr[0][n, 4, 9]) < 1e-4
    assert out[0 * 9 + 4, 5, 0] == 0     # dx=-4 at x=0 falls outside -> zero padding


@pytest.mark.parametrize('which', [2, 5, 9])
def test_tracker_real_128(real_weights, which):
    g = golden('track_real_128.npz')
    frames = g['frames']
    trk = O.OracleTracker(real_weights, deltas=g['deltas'].tolist())
    trk.init(frames[0])
    for i in range(1, which + 1):
        m = trk.track(frames[i])
        got = np.concatenate(m.result, 0)
        assert np.abs(got.reshape(4, -1).mean(1) - g['means'][i - 1]).max() < 2e-3
    ref = g[f'result_{which}']
    bad = (np.abs(got - ref) > 2e-3).any(0)
    assert bad.mean() < 0.01, bad.mean()


def test_warp_forward_matches_reference():
    """Forward splat: oracle restatement of FlowOUTrackingResult.warp_forward / bilinear_splat against the
    vectors recorded from the unmodified reference (clamped corner indices, mask, border)."""
    g = golden('warp_forward.npz')
    assert np.abs(O.warp_forward(g['flow'], g['img']) - g['out_plain']).max() < 1e-6
    assert np.abs(O.warp_forward(g['flow'], g['img'], g['mask'], -1.0) - g['out_mask']).max() < 1e-6


def test_point_queries_match_reference():
    """Point queries: the oracle's bilinear sampler / chain against FlowOUTrackingResult.sample /
    warp_forward_points / chain / warp_backward of the unmodified reference, incl. queries on the border and outside."""
    g = golden('point_queries.npz')
    packed = np.concatenate([g['flow'], g['occlusion'], g['sigma']])
    q = g['queries']
    s = O.bilinear_zero(packed, q[:, 0], q[:, 1], via_mul=True)
    assert np.abs(s[:2] - g['sample_flow']).max() < 1e-4
    assert np.abs(s[2:3] - g['sample_occlusion']).max() < 1e-5 and np.abs(s[3:4] - g['sample_sigma']).max() < 1e-5
    assert np.abs((q + s[:2].T) - g['warped_points']).max() < 1e-4
    assert np.abs(s[2] - g['pt_occlusion']).max() < 1e-5 and np.abs((q + s[:2].T) - g['pt_coords']).max() < 1e-4
    z = np.zeros_like(g['occlusion'])
    wf, _, _ = O.chain((g['flow'], g['occlusion'], g['sigma']), (g['other'], z, z))
    assert np.abs(wf - g['chained']).max() < 1e-4
