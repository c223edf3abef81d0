"""Copy the reference's DATA inputs (checkpoint and demo video -- never sources) into oracle/_ref/.

oracle/_ref/ is git-ignored but not gpurun-ignored, so the shipped RAFT-OU weights travel to the
GPU box where /root/reference does not exist.  Run from ``__graft_entry__.build()`` when the
reference checkout is present.  TEST INFRASTRUCTURE ONLY.
"""
import os
import shutil

from . import ref_bridge as R

DST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
CKPT_DST = os.path.join(DST_DIR, 'raft_ou_checkpoint.pth')
VIDEO_DST = os.path.join(DST_DIR, 'demo_video.mp4')


def fetch():
    src = os.path.join(R.REF_ROOT, R.CKPT_REL)
    if not os.path.isfile(src):
        return None
    os.makedirs(DST_DIR, exist_ok=True)
    if not os.path.isfile(CKPT_DST) or os.path.getsize(CKPT_DST) != os.path.getsize(src):
        shutil.copyfile(src, CKPT_DST)
    vsrc = os.path.join(R.REF_ROOT, R.VIDEO_REL)
    if os.path.isfile(vsrc) and (not os.path.isfile(VIDEO_DST) or os.path.getsize(VIDEO_DST) != os.path.getsize(vsrc)):
        shutil.copyfile(vsrc, VIDEO_DST)
    return CKPT_DST


def demo_frames(n, size=None):
    """First n frames of the demo video (the travelled copy, else the reference checkout), resized like SURVEY 8d configs 1-2."""
    import cv2
    import numpy as np
    path = next((p for p in (VIDEO_DST, os.path.join(R.REF_ROOT, R.VIDEO_REL)) if os.path.isfile(p)), None)
    if path is None:
        return []
    cap = cv2.VideoCapture(path)
    frames = []
    while len(frames) < n:
        ok, f = cap.read()
        if not ok:
            break
        if size is not None:
            f = cv2.resize(f, size, interpolation=cv2.INTER_AREA)
        frames.append(np.ascontiguousarray(f))
    cap.release()
    return frames


def find_checkpoint():
    """Shipped checkpoint if reachable (travelled copy first), else None."""
    for p in (CKPT_DST, os.path.join(R.REF_ROOT, R.CKPT_REL)):
        if os.path.isfile(p):
            return p
    return None


if __name__ == '__main__':
    print(fetch())
