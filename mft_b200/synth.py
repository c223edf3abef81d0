"""Seeded synthetic video for benchmarks and full-size tests: a band-limited random
texture seen through a smooth time-varying affine + low-frequency displacement field, with a few
independently moving textured rectangles that create real occlusions.  Pure uniform noise is
deliberately avoided (no trackable structure)."""
import cv2
import numpy as np


def _texture(rng, h, w):
    acc = np.zeros((h, w, 3), np.float32)
    for sigma, gain in ((1.5, 0.5), (4.0, 0.8), (12.0, 1.0)):
        n = rng.uniform(-1, 1, (h, w, 3)).astype(np.float32)
        n = cv2.GaussianBlur(n, (0, 0), sigma)
        acc += gain * n / (n.std() + 1e-6)
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return (acc * 255).astype(np.uint8)


def synthetic_video(T, H, W, seed=1234, max_motion=2.5, n_rects=3):
    """Generator of T uint8 BGR frames (H,W,3)."""
    rng = np.random.default_rng(seed)
    m = 96
    tex = _texture(rng, H + 2 * m, W + 2 * m)
    disp = rng.uniform(-1, 1, (2, H // 32 + 2, W // 32 + 2)).astype(np.float32)
    disp = np.stack([cv2.resize(d, (W, H), interpolation=cv2.INTER_CUBIC) for d in disp])
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    cx, cy = W / 2.0, H / 2.0
    rects = []
    for _ in range(n_rects):
        rh, rw = int(rng.integers(H // 10, H // 4)), int(rng.integers(W // 10, W // 4))
        rects.append(dict(tex=_texture(rng, rh, rw), p=np.array([rng.uniform(0, W - rw), rng.uniform(0, H - rh)]),
                          v=rng.uniform(-max_motion, max_motion, 2)))
    phase = rng.uniform(0, 2 * np.pi, 4)
    for t in range(T):
        ang = 0.06 * np.sin(2 * np.pi * t / 97.0 + phase[0])
        sc = 1.0 + 0.04 * np.sin(2 * np.pi * t / 61.0 + phase[1])
        tx = 0.6 * max_motion * t * np.cos(phase[2]) % (m / 2)
        ty = 0.6 * max_motion * t * np.sin(phase[2]) % (m / 2)
        amp = 3.0 * np.sin(2 * np.pi * t / 45.0 + phase[3])
        ca, sa = np.cos(ang) * sc, np.sin(ang) * sc
        mx = ca * (xs - cx) - sa * (ys - cy) + cx + m + tx + amp * disp[0]
        my = sa * (xs - cx) + ca * (ys - cy) + cy + m + ty + amp * disp[1]
        frame = cv2.remap(tex, mx.astype(np.float32), my.astype(np.float32), cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        for r in rects:
            rh, rw = r['tex'].shape[:2]
            x0 = int(round((r['p'][0] + r['v'][0] * t) % (W - rw)))
            y0 = int(round((r['p'][1] + r['v'][1] * t) % (H - rh)))
            frame[y0:y0 + rh, x0:x0 + rw] = r['tex']
        yield np.ascontiguousarray(frame)


DEMO_VIDEO_REL = 'demo_in/ugsJtsO9w1A-00.00.24.457-00.00.29.462_HD.mp4'       # the reference's demo input (demo.py)


def find_demo_video():
    """Path of the reference's demo video if it is reachable (it is DATA: $MFT_DEMO_VIDEO, the copy that travels with the
    repo snapshot, a reference checkout); None otherwise."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.environ.get('MFT_DEMO_VIDEO', ''), os.path.join(root, 'oracle', '_ref', 'demo_video.mp4'),
              os.path.join(os.environ.get('MFT_REFERENCE_ROOT', '/root/reference'), DEMO_VIDEO_REL)):
        if p and os.path.isfile(p):
            return p
    return None


def demo_video_frames(size=(512, 512), max_frames=None):
    """All frames of the demo video as uint8 BGR, resized with cv2.INTER_AREA to (W, H) = size (BASELINE config 2)."""
    path = find_demo_video()
    if path is None:
        return []
    cap = cv2.VideoCapture(path)
    frames = []
    while max_frames is None or len(frames) < max_frames:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(np.ascontiguousarray(cv2.resize(f, size, interpolation=cv2.INTER_AREA)))
    cap.release()
    return frames
