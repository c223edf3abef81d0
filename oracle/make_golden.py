"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference ships no golden vectors of its own, so these files ARE the pin
for oracle/mft_oracle.py.  Deterministic for a fixed thread count (4 here).

Files
  raft_real_256.npz     config 1 of BASELINE.json: demo frames 0,1 at 256x256, shipped checkpoint,
                        12 iters -> flow / occlusion / sigma
  raft_seeded_128.npz   seeded stand-in weights (oracle.seeded_weights(0)), demo frames 0,1,8 at
                        128x128: fnet/cnet outputs of frame 0 and final outputs of pairs (0,1),(0,8)
  raft_real_128.npz     same pairs with the shipped checkpoint (needs the checkpoint at test time)
  chain_select.npz      MFT.track's chaining + selection + invalid mask through the reference's own
                        tracker class around a replay flower (synthetic flows incl. overflow/ties/all-occluded/OOB)
  raft_seeded_pad.npz   seeded weights, demo frames 0,2 at 131x140 (replicate pad 2|3 rows, 2|2 columns, then unpad)
  warp_forward.npz      FlowOUTrackingResult.warp_forward (bilinear forward splat, results.py:190-248) of a random image
                        through a random flow with end points outside the grid, with and without mask / border
  point_queries.npz     FlowOUTrackingResult.chain / warp_backward / warp_forward_points / sample / invalid_mask and
                        point_tracking.convert_to_point_tracking on a random result, queries inside / on the border / outside
  track_real_128.npz    10 demo frames at 128x128, deltas [inf,1,2,4,8], shipped checkpoint: the
                        reference tracker's per-frame results (full field for 3 frames + per-frame sums)
"""
from __future__ import annotations

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


def raft_pairs(model, frames, pairs, tag):
    fl = R.CpuFlower(model)
    out = {}
    for a, b in pairs:
        flow, ex = fl.compute_flow(frames[a], frames[b])
        out[f'flow_{a}_{b}'] = _np(flow)
        out[f'occ_{a}_{b}'] = _np(ex['occlusion'])
        out[f'sigma_{a}_{b}'] = _np(ex['sigma'])
        out[f'coords_{a}_{b}'] = _np(ex['raw']['coords'][0])
    return out


def warp_forward_case():
    R._import_reference()
    from MFT.results import FlowOUTrackingResult as RefRes
    rng = np.random.default_rng(77)
    H, W = 40, 56
    flow = (rng.standard_normal((2, H, W)) * 6).astype(np.float32)
    flow[:, :4, :] += 30
    flow[:, -3:, :] -= 40          # end points outside the grid on both sides
    img = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    mask = rng.uniform(0, 1, (H, W)) > 0.3
    res = RefRes(torch.from_numpy(flow), torch.zeros(1, H, W), torch.zeros(1, H, W))
    return dict(flow=flow, img=img, mask=mask, out_plain=res.warp_forward(img).astype(np.float32),
                out_mask=res.warp_forward(img, mask=mask, border=-1.0).astype(np.float32))


def padded_case():
    """One pair at 131x140 (pad 5 -> 2|3 rows, 4 -> 2|2 columns: InputPadder, core/utils/utils.py:9-24) through the reference's
    compute_flow with the seeded weights: pins replicate padding + unpadding, incl. the odd split."""
    import cv2
    Wseed = O.seeded_weights(0)
    seeded = R.build_reference_model(Wseed)
    fr = [cv2.resize(f, (140, 131), interpolation=cv2.INTER_AREA) for f in R.demo_frames(3, size=(256, 256))]
    g = raft_pairs(seeded, {0: fr[0], 2: fr[2]}, [(0, 2)], 'pad')
    g['frames'] = np.stack([fr[0], fr[2]])
    return g


def point_query_case():
    """FlowOUTrackingResult geometry + point_tracking.convert_to_point_tracking of the unmodified reference (results.py:87-188,
    250-265; point_tracking.py:6-27) on a random result: queries on pixel centres, between them, on the border and outside."""
    R._import_reference()
    from MFT.results import FlowOUTrackingResult as RefRes
    from MFT.point_tracking import convert_to_point_tracking
    rng = np.random.default_rng(99)
    H, W = 36, 52
    flow = (rng.standard_normal((2, H, W)) * 5).astype(np.float32)
    flow[:, :3, :] -= 20
    flow[:, :, -4:] += 25           # end points outside the image
    occ = rng.uniform(0, 1, (1, H, W)).astype(np.float32)
    sigma = rng.uniform(0.1, 4, (1, H, W)).astype(np.float32)
    other = (rng.standard_normal((2, H, W)) * 3).astype(np.float32)
    img = rng.uniform(0, 1, (3, H, W)).astype(np.float32)
    q = np.concatenate([rng.uniform([0, 0], [W - 1, H - 1], (40, 2)),
                        np.array([[0, 0], [W - 1, H - 1], [W - 1, 0], [0, H - 1], [7, 11], [W - 1.5, H - 1.25]]),
                        np.array([[-0.5, 3.0], [W - 0.5, 3.0], [5.0, -0.75], [5.0, H - 0.25], [-3.0, -3.0], [W + 4.0, H + 9.0]]),
                        ]).astype(np.float32)
    res = RefRes(torch.from_numpy(flow), torch.from_numpy(occ), torch.from_numpy(sigma))
    qt = torch.from_numpy(q)
    sf, so, ss = res.sample(qt)
    pc, po = convert_to_point_tracking(res, qt)
    return dict(flow=flow, occlusion=occ, sigma=sigma, other=other, img=img, queries=q,
                warped_points=_np(res.warp_forward_points(qt)), sample_flow=_np(sf), sample_occlusion=_np(so),
                sample_sigma=_np(ss), pt_coords=np.asarray(pc, np.float32), pt_occlusion=np.asarray(po, np.float32),
                chained=_np(res.chain(torch.from_numpy(other))), warped_img=_np(res.warp_backward(torch.from_numpy(img))),
                invalid=res.invalid_mask().numpy().astype(np.bool_))


def chain_select_case(rng, H, W, K, thr):
    """Synthetic per-chain inputs that exercise the selection edge cases."""
    lefts, rights = [], []
    for k in range(K):
        lf = (rng.standard_normal((2, H, W)) * 5).astype(np.float32)
        lo = rng.uniform(0, 0.04, (1, H, W)).astype(np.float32)
        ls = rng.uniform(0.0, 2.0, (1, H, W)).astype(np.float32)
        rf = (rng.standard_normal((2, H, W)) * 3).astype(np.float32)
        ro = rng.uniform(0, 0.03, (1, H, W)).astype(np.float32)
        rs = rng.uniform(0.05, 2.0, (1, H, W)).astype(np.float32)
        lefts.append([lf, lo, ls])
        rights.append([rf, ro, rs])
    # region A: everything occluded -> index 0
    for k in range(K):
        lefts[k][1][:, :4, :] = 0.5
    # region B: exact ties between all candidates -> lowest index.  Zero left flow keeps every
    # bilinear tap of rows 5..6 inside rows 4..7, where all candidates are made identical.
    lefts[0][0][:, 4:8, :] = 0
    for k in range(1, K):
        for j in range(3):
            lefts[k][j][:, 4:8, :] = lefts[0][j][:, 4:8, :]
            rights[k][j][:, 4:8, :] = rights[0][j][:, 4:8, :]
    # region C: flows pointing far outside the image
    lefts[0][0][:, 8:10, :] = 1000.0
    lefts[1 % K][0][:, 9:10, :] = -1000.0
    # region D: integer flows (sampling exactly on pixel centres, incl. the last row/column)
    for k in range(K):
        lefts[k][0][:, 10:14, :] = np.round(lefts[k][0][:, 10:14, :])
    # region E: huge sigma on one candidate (its square overflows to +inf in the chain).  NaN /
    # inf inputs cannot reach the reference's selection: bilinear weights turn inf into NaN and
    # the FlowOUTrackingResult constructor asserts sigma >= 0 (results.py:31-33)
    rights[1 % K][2][:, 14, W // 2:] = 1e30
    return lefts, rights


def run_reference_select(lefts, rights, thr):
    """Drive the reference's MFT.track with recorded flows: every chain k uses delta k+1's
    slot; memory is pre-seeded with the 'left' results.  Returns the tracker's result."""
    ref_mft, ref_results, ref_config, _, _ = R._import_reference()
    FR = ref_results.FlowOUTrackingResult
    K = len(lefts)
    H, W = lefts[0][0].shape[1:]

    class Replay:
        def __init__(self):
            self.calls = []

        def compute_flow(self, left_img, right_img, mode='flow', init_flow=None):
            k = int(left_img[0, 0, 0])
            self.calls.append(k)
            f, o, s = (torch.from_numpy(a.copy()) for a in rights[k])
            return f, {'occlusion': o, 'sigma': s}

    class CpuMFT(ref_mft.MFT):
        def __init__(self, config, flower):
            self.C, self.flower, self.device = config, flower, 'cpu'

    C = ref_config.Config()
    # chain order after MFT.py:114's sort: inf first then ascending -> chain k <-> k-th entry
    C.deltas = [np.inf] + list(range(1, K))
    C.occlusion_threshold = thr
    trk = CpuMFT(C, Replay())
    img0 = np.zeros((H, W, 3), np.uint8)
    trk.init(img0, start_frame_i=0)
    trk.current_frame_i = K            # next track() is frame K+1 ... arrange left ids
    cur = K + 1
    # left_id for inf = 0 -> chain 0 ; left_id for delta d = cur-d -> chain d
    trk.memory = {}
    for k in range(K):
        left_id = 0 if k == 0 else cur - k
        img = np.full((H, W, 3), k, np.uint8)
        trk.memory[left_id] = {'img': img, 'result': FR(*(torch.from_numpy(a.copy()) for a in lefts[k]))}
    with np.errstate(all='ignore'):
        meta = trk.track(np.zeros((H, W, 3), np.uint8))
    r = meta.result
    return _np(r.flow), _np(r.occlusion), _np(r.sigma)


def main():
    warnings.filterwarnings('ignore')
    assert R.available(), 'reference checkout not found'
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)

    real = R.build_reference_model()
    # --- config 1: 256x256, frames 0,1 ---------------------------------------------------------
    fr256 = R.demo_frames(2, size=(256, 256))
    g = raft_pairs(real, fr256, [(0, 1)], 'real256')
    g['frames'] = np.stack(fr256)
    np.savez_compressed(os.path.join(OUT, 'raft_real_256.npz'), **g)

    # --- 128x128: seeded and real weights ------------------------------------------------------
    fr128 = R.demo_frames(10, size=(128, 128))
    sel = {0: fr128[0], 1: fr128[1], 8: fr128[8]}
    Wseed = O.seeded_weights(0)
    seeded = R.build_reference_model(Wseed)
    for tag, model in (('seeded', seeded), ('real', real)):
        g = raft_pairs(model, sel, [(0, 1), (0, 8)], tag)
        x = O.bgr_to_input(sel[0])
        x = 2 * (x / 255.0) - 1.0
        with torch.no_grad():
            g['fnet_0'] = _np(model.fnet(x)[0])
            g['cnet_0'] = _np(model.cnet(x)[0])
        g['frames'] = np.stack([sel[0], sel[1], sel[8]])
        g['frame_ids'] = np.array([0, 1, 8])
        np.savez_compressed(os.path.join(OUT, f'raft_{tag}_128.npz'), **g)

    # --- chain + select through the reference tracker ----------------------------------------
    rng = np.random.default_rng(20260101)
    g = {}
    ncase = 0
    for (H, W, K) in ((24, 40, 7), (17, 33, 3), (16, 16, 1), (32, 24, 5)):
        lefts, rights = chain_select_case(rng, H, W, K, 0.02)
        res = run_reference_select(lefts, rights, 0.02)
        assert res is not None
        g[f'c{ncase}_left'] = np.stack([np.concatenate(l, 0) for l in lefts])     # (K,4,H,W)
        g[f'c{ncase}_right'] = np.stack([np.concatenate(r, 0) for r in rights])
        g[f'c{ncase}_out'] = np.concatenate(res, 0)                                 # (4,H,W)
        ncase += 1
    g['ncase'] = np.array(ncase)
    g['thr'] = np.array(0.02, np.float32)
    np.savez_compressed(os.path.join(OUT, 'chain_select.npz'), **g)
    np.savez_compressed(os.path.join(OUT, 'warp_forward.npz'), **warp_forward_case())
    np.savez_compressed(os.path.join(OUT, 'point_queries.npz'), **point_query_case())
    np.savez_compressed(os.path.join(OUT, 'raft_seeded_pad.npz'), **padded_case())

    # --- short real tracking run ---------------------------------------------------------------
    deltas = [np.inf, 1, 2, 4, 8]
    trk = R.build_reference_tracker(real, deltas)
    trk.init(fr128[0])
    g = {'frames': np.stack(fr128), 'deltas': np.array(deltas)}
    sums = []
    for i in range(1, len(fr128)):
        r = trk.track(fr128[i]).result
        full = np.concatenate([_np(r.flow), _np(r.occlusion), _np(r.sigma)], 0)
        sums.append(full.reshape(4, -1).astype(np.float64).mean(1))
        if i in (2, 5, 9):
            g[f'result_{i}'] = full
    g['means'] = np.stack(sums)
    np.savez_compressed(os.path.join(OUT, 'track_real_128.npz'), **g)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
