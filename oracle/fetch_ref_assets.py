"""Copy the reference's DATA inputs (checkpoint only -- never sources) into oracle/_ref/.

oracle/_ref/ is git-ignored but not gpurun-ignored, so the shipped RAFT-OU weights travel to the
GPU box where /root/reference does not exist.  Run from ``__graft_entry__.build()`` when the
reference checkout is present.  TEST INFRASTRUCTURE ONLY.
"""
import os
import shutil

from . import ref_bridge as R

DST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
CKPT_DST = os.path.join(DST_DIR, 'raft_ou_checkpoint.pth')


def fetch():
    src = os.path.join(R.REF_ROOT, R.CKPT_REL)
    if not os.path.isfile(src):
        return None
    os.makedirs(DST_DIR, exist_ok=True)
    if not os.path.isfile(CKPT_DST) or os.path.getsize(CKPT_DST) != os.path.getsize(src):
        shutil.copyfile(src, CKPT_DST)
    return CKPT_DST


def find_checkpoint():
    """Shipped checkpoint if reachable (travelled copy first), else None."""
    for p in (CKPT_DST, os.path.join(R.REF_ROOT, R.CKPT_REL)):
        if os.path.isfile(p):
            return p
    return None


if __name__ == '__main__':
    print(fetch())
