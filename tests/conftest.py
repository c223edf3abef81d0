import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope='session')
def real_weights():
    """The shipped RAFT-OU checkpoint through the oracle's loader (skip if it did not travel)."""
    from oracle import fetch_ref_assets, mft_oracle
    path = fetch_ref_assets.find_checkpoint()
    if path is None:
        pytest.skip('shipped checkpoint not available on this box')
    return mft_oracle.load_checkpoint(path)


@pytest.fixture(scope='session')
def seeded_weights():
    from oracle import mft_oracle
    return mft_oracle.seeded_weights(0)


# ---- measured parity numbers ------------------------------------------------------------------------------------------
# Every parity test reports the numbers it measured (not only pass / fail): they are printed in the test's captured
# output and collected into gpurun_out/parity_measured.json, from which the gates in the tests were set (<= 3x measured).
_PARITY = {}


def record_parity(name, stats):
    clean = {k: (float(v) if isinstance(v, (int, float, np.floating, np.integer)) else v) for k, v in stats.items()}
    _PARITY[name] = clean
    print(f'[parity] {name}: ' + ', '.join(f'{k}={v:.6g}' if isinstance(v, float) else f'{k}={v}' for k, v in clean.items()))


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_measured.json'), 'w') as f:
            json.dump(_PARITY, f, indent=1, sort_keys=True)
    except OSError:
        pass


REFDATA = os.path.join(ROOT, 'oracle', '_ref')


def frames_for(golden_npz, name, regenerate=None):
    """The uint8 input frames a full-size golden file was recorded on: regenerated (synthetic video) when that
    reproduces the recorded CRC32s bit for bit, else the copy that travelled in oracle/_ref/ (git-ignored data, like
    the checkpoint); skips when neither is available."""
    import zlib
    want = [int(c) for c in golden_npz['frame_crc']]

    def ok(fr):
        return len(fr) >= len(want) and all((zlib.crc32(np.ascontiguousarray(f).tobytes()) & 0xffffffff) == c for f, c in zip(fr, want))
    if regenerate is not None:
        fr = regenerate()
        if ok(fr):
            return list(fr[:len(want)]), 'regenerated'
    path = os.path.join(REFDATA, f'frames_{name}.npy')
    if os.path.isfile(path):
        fr = np.load(path)
        if ok(fr):
            return [np.ascontiguousarray(f) for f in fr[:len(want)]], 'oracle/_ref'
    pytest.skip(f'input frames of {name} not reproducible here and oracle/_ref/frames_{name}.npy did not travel')
