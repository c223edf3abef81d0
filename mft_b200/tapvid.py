"""TAP-Vid style evaluation loop over the tracker: the caller on the far side of the hot path.

Mirrors the reference's runner for this path -- MFT/runners/run_MFT_tapvid.py:116-285 (per sequence: one flow cache, per
query mode and start frame a forward (+ backward for 'strided') ``track_sequence``, every frame's result turned into
point tracks by ``convert_to_point_tracking``) and the dataset schema of MFT/evaluation/tapvid_eval_stuff.py:636-665
(``{name: {'video': uint8 (T,H,W,3) RGB, 'points': float (n,T,2) xy in [0,1], 'occluded': bool (n,T)}}``, videos
resized 256x256 -> 512x512 as dataset_configs/pkl-tapvid-davis-256x256_512x512.py asks).  Differences, all on the
device side of the boundary: results stay in HBM (``track(..., device_result=True)``), queries are sampled by one
kernel launch per frame and only (n,2)+(n,) floats reach the host, the flow cache is ``DeviceFlowCache``.

The TAP-Vid pickles are not redistributable with this repo; ``synthetic_dataset`` builds one with the same schema
(point tracks are plausible but NOT ground truth: this module measures throughput and exercises the call sequence,
accuracy evaluation stays with the reference's evaluation code).
"""
import numpy as np
import torch

from .point_tracking import convert_to_point_tracking
from .synth import synthetic_video


def synthetic_dataset(n_sequences=8, n_frames=24, n_points=32, size=256, seed=1234):
    """{name: {'video', 'points', 'occluded'}} with the TAP-Vid schema (tapvid_eval_stuff.py:636-665)."""
    rng = np.random.default_rng(seed)
    data = {}
    for i in range(n_sequences):
        video = np.stack([f[:, :, ::-1] for f in synthetic_video(n_frames, size, size, seed=seed + i)])       # RGB like the pickles
        p0 = rng.uniform(0.1, 0.9, (n_points, 1, 2))
        walk = np.cumsum(rng.normal(0, 0.004, (n_points, n_frames, 2)), axis=1)
        occluded = rng.uniform(0, 1, (n_points, n_frames)) < 0.15
        occluded[np.arange(n_points), rng.integers(0, max(1, n_frames // 3), n_points)] = False      # every point is visible somewhere early
        data[f'synth-{i:03d}'] = {'video': np.ascontiguousarray(video), 'points': np.clip(p0 + walk, 0.0, 1.0), 'occluded': occluded}
    return data


def resize_video(video, hw):
    """(T,H,W,3) uint8 -> (T,h,w,3) (tapvid_eval_stuff.py:61-79 uses mediapy.resize_video; cv2 here, host side)."""
    import cv2
    h, w = hw
    if video.shape[1:3] == (h, w):
        return video
    return np.stack([cv2.resize(f, (w, h), interpolation=cv2.INTER_LINEAR) for f in video])


def sample_queries_first(occluded, points):
    """(t, y, x) of every point at its first visible frame (TAP-Vid 'first' mode)."""
    q = []
    for i in range(points.shape[0]):
        vis = np.where(~occluded[i])[0]
        if len(vis):
            t = int(vis[0])
            q.append((t, points[i, t, 1], points[i, t, 0]))
    return np.asarray(q, np.float64).reshape(-1, 3)


def sample_queries_strided(occluded, points, stride=5):
    """(t, y, x) of every point visible at frames 0, stride, 2*stride, ... (TAP-Vid 'strided' mode)."""
    q = []
    for t in range(0, points.shape[1], stride):
        for i in np.where(~occluded[:, t])[0]:
            q.append((t, points[i, t, 1], points[i, t, 0]))
    return np.asarray(q, np.float64).reshape(-1, 3)


def track_sequence(tracker, video, start_frame, direction='forward', debug=False, flow_cache=None, device_result=True,
                   on_frame=None):
    """run_MFT_tapvid.py:251-285: init at ``start_frame``, then track to the end (or back to frame 0).  video: (T,H,W,3)
    uint8 BGR (numpy, or a CUDA tensor: frames already in HBM).  on_frame(frame_i, meta) is called per frame instead of
    keeping every meta alive (the reference keeps them all: T x 4 MiB of host memory per run)."""
    assert direction in ('forward', 'backward')
    n = video.shape[0]
    frames = range(start_frame, n) if direction == 'forward' else range(start_frame, -1, -1)
    metas = {}
    for k, frame_i in enumerate(frames):
        frame = video[frame_i]
        if k == 0:
            meta = tracker.init(frame, start_frame_i=start_frame, time_direction=+1 if direction == 'forward' else -1,
                                flow_cache=flow_cache)
        else:
            try:
                meta = tracker.track(frame, debug=debug, device_result=device_result)
            except StopIteration:
                break
        meta.frame_i = frame_i
        meta.backward = direction == 'backward'
        if on_frame is not None:
            on_frame(frame_i, meta)
        else:
            metas[frame_i] = meta
    return metas


def run_sequence(tracker, video_bgr, query_points, query_mode, flow_cache=None, device='cuda'):
    """One (sequence, query mode) of run_MFT_tapvid.py:140-238.  query_points: (n, 3) (t, y, x) in pixels of the video.
    Returns (pred_tracks (n,T,2) xy, pred_occluded (n,T), frames passed to the tracker)."""
    n_frames = video_bgr.shape[0]
    qp = np.asarray(query_points).astype(np.int64)
    pred_tracks = np.zeros((qp.shape[0], n_frames, 2))
    pred_occluded = np.zeros((qp.shape[0], n_frames))
    n_tracked = 0
    for start_frame in np.unique(qp[:, 0]):
        mask = qp[:, 0] == start_frame
        queries = torch.from_numpy(qp[mask, 1:][:, ::-1].copy()).to(device)                  # xy order
        for direction in (['forward', 'backward'] if query_mode == 'strided' else ['forward']):
            def on_frame(frame_i, meta):
                coords, occl = convert_to_point_tracking(meta.result, queries)
                pred_tracks[mask, frame_i, :] = coords
                pred_occluded[mask, frame_i] = occl
            track_sequence(tracker, video_bgr, int(start_frame), direction, flow_cache=flow_cache, on_frame=on_frame)
            n_tracked += (n_frames - start_frame) if direction == 'forward' else (start_frame + 1)
    return pred_tracks, pred_occluded, n_tracked
