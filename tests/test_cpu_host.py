"""Host-side logic on CPU: C-ABI surface, weight packing, config semantics, tracker bookkeeping,
world-size-2 gloo run of the flow-sharding logic.  No GPU, no compute calls into the library."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from mft_b200 import _lib
    import glob
    headers = sorted(glob.glob(os.path.join(ROOT, 'include', '*.h'))) + [os.path.join(ROOT, 'mft_b200', 'csrc', 'mft_b200_internal.h')]
    header = ''.join(open(h).read() for h in headers)
    public = open(os.path.join(ROOT, 'include', 'mft_b200.h')).read()
    for hook in ('conv2d_test', 'conv2d_bench', 'set_global_option', 'debug_read'):     # tuning hooks stay out of the boundary
        assert 'mftb200_' + hook not in public
    declared = sorted(set(re.findall(r'\b(mftb200_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 20
    bound = set(_lib.exported_symbols())                  # dlopen + getattr of every bound symbol
    nm = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf'\bT {name}\b', nm), f'{name} declared in the header but not exported'
        assert name in bound, f'{name} exported but not bound in mft_b200/_lib.py'
    assert _lib.lib().mftb200_version().decode().startswith('mft_b200')


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'mft_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f


def _unpack(packed, cout):
    w16, bias, cout_pad, ktot, bias_len = packed
    return torch.from_numpy(w16.view(np.float16).astype(np.float32))[:cout], torch.from_numpy(bias)[:cout]


def _im2col(x, kh, kw, cin_pad):
    """x (1,C,H,W) -> (H*W, taps*cin_pad) with K order (tap, channel), zero padded."""
    _, C, H, W = x.shape
    xp = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2))
    cols = []
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[0, :, ky:ky + H, kx:kx + W].reshape(C, H * W).t()
            cols.append(F.pad(patch, (0, cin_pad - C)))
    return torch.cat(cols, 1)


@pytest.mark.parametrize('shape', [(24, 7, 3, 3), (130, 20, 1, 5), (70, 33, 5, 1), (324, 16, 1, 1)])
def test_pack_layout_matches_conv(shape):
    from mft_b200 import weights as WT
    cin, cout, kh, kw = shape
    g = torch.Generator().manual_seed(0)
    w = torch.randn(cout, cin, kh, kw, generator=g).half().float()
    b = torch.randn(cout, generator=g)
    x = torch.randn(1, cin, 6, 9, generator=g)
    W2, b2 = _unpack(WT._pack(w, b), cout)
    cin_pad = (cin + 63) // 64 * 64
    got = _im2col(x, kh, kw, cin_pad) @ W2.t() + b2
    ref = F.conv2d(x, w, b, padding=(kh // 2, kw // 2))[0].reshape(cout, -1).t()
    assert (got - ref).abs().max() < 1e-3


def test_pack_all_semantics(seeded_weights):
    """BN folding, z|r stacking, q input permutation, OU stacking / block diagonal, 7x7 im2col order."""
    from mft_b200 import weights as WT
    W = seeded_weights
    P = dict(zip(WT.LAYER_NAMES, WT.pack_all(W)))
    g = torch.Generator().manual_seed(1)
    # cnet layer1.0.conv1 + BN folded == oracle conv -> batch_norm_eval
    x = torch.randn(1, 64, 5, 7, generator=g)
    ref = O.batch_norm_eval(O._conv(x, W, 'cnet.layer1.0.conv1', padding=1), W, 'cnet.layer1.0.norm1')
    W2, b2 = _unpack(P['cnet.layer1.0.conv1'], 64)
    got = (_im2col(x, 3, 3, 64) @ W2.t() + b2).t().reshape(1, 64, 5, 7)
    assert (got - ref).abs().max() < 5e-3
    # downsample conv folds norm3
    ref = O.batch_norm_eval(O._conv(x, W, 'cnet.layer2.0.downsample.0'), W, 'cnet.layer2.0.norm3')
    W2, b2 = _unpack(P['cnet.layer2.0.downsample.0'], 96)
    assert ((_im2col(x, 1, 1, 64) @ W2.t() + b2).t().reshape(1, 96, 5, 7) - ref).abs().max() < 5e-3
    # GRU: z|r stacked; q consumes [inp | motion | r*h]
    h, inp, mot, rh = (torch.randn(1, 128, 4, 6, generator=g) for _ in range(4))
    hx = torch.cat([h, inp, mot], 1)
    W2, b2 = _unpack(P['gru_zr1'], 256)
    got = (_im2col(hx, 1, 5, 384) @ W2.t() + b2).t().reshape(1, 256, 4, 6)
    assert (got[:, :128] - O._conv(hx, W, 'update_block.gru.convz1', padding=(0, 2))).abs().max() < 5e-3
    assert (got[:, 128:] - O._conv(hx, W, 'update_block.gru.convr1', padding=(0, 2))).abs().max() < 5e-3
    W2, b2 = _unpack(P['gru_q2'], 128)
    got = (_im2col(torch.cat([inp, mot, rh], 1), 5, 1, 384) @ W2.t() + b2).t().reshape(1, 128, 4, 6)
    ref = O._conv(torch.cat([rh, inp, mot], 1), W, 'update_block.gru.convq2', padding=(2, 0))
    assert (got - ref).abs().max() < 5e-3
    # OU heads: conv1 stacked, conv2 block diagonal
    x712 = torch.randn(1, 712, 4, 6, generator=g)
    W2, b2 = _unpack(P['ou1'], 256)
    hid = torch.relu((_im2col(x712, 3, 3, 768) @ W2.t() + b2).t().reshape(1, 256, 4, 6))
    W3, b3 = _unpack(P['ou2'], 3)
    got = (_im2col(hid, 3, 3, 256) @ W3.t() + b3).t().reshape(1, 3, 4, 6)
    occ = O._conv(torch.relu(O._conv(x712, W, 'occlusion_block.occl_head.conv1', padding=1)), W, 'occlusion_block.occl_head.conv2', padding=1)
    unc = O._conv(torch.relu(O._conv(x712, W, 'occlusion_block.uncertainty_head.conv1', padding=1)), W,
                  'occlusion_block.uncertainty_head.conv2', padding=1)
    assert (got[:, :2] - occ).abs().max() < 2e-2 and (got[:, 2:] - unc).abs().max() < 2e-2
    # 7x7 convs are stored for im2col'ed operands: k = (ky*7+kx)*cin + c
    flow = torch.randn(1, 2, 9, 9, generator=g)
    W2, b2 = _unpack(P['convf1'], 128)
    patch = F.pad(flow, (3, 3, 3, 3))[0, :, 4:11, 2:9].permute(1, 2, 0).reshape(1, 98)      # window around (y=4+3.., x=2+3..)
    ref = O._conv(flow, W, 'update_block.encoder.convf1', padding=3)[0, :, 4, 2]
    assert ((F.pad(patch, (0, 30)) @ W2.t() + b2)[0] - ref).abs().max() < 5e-3


def test_config_semantics(tmp_path):
    from mft_b200.config import Config, load_config
    C = Config()
    assert not C.timers_enabled and not C.foo.bar.baz          # missing -> falsy empty Config
    C.deltas = [1, 2]
    assert C.deltas == [1, 2]
    p = tmp_path / 'cfg.py'
    p.write_text('from mft_b200.config import Config\ndef get_config():\n    c = Config(); c.x = 3; return c\n')
    assert load_config(p).x == 3
    with pytest.raises(AssertionError):
        load_config(tmp_path / 'missing.py')


class _FakeEngine:
    def __init__(self):
        self.encoded = {}

    def encode_frame(self, img, slot):
        self.encoded[slot] = int(img[0, 0, 0])
        return False

    def error_flag_async(self):
        pass

    def error_flag_poll(self):
        pass

    def wait_frame_copied(self):
        pass


def _tracker_without_gpu(deltas, direction, start, monkeypatch):
    """mft_b200.MFT.MFT with the engine calls replaced, to exercise the host bookkeeping only."""
    import mft_b200.MFT as M
    from mft_b200.config import Config
    from mft_b200.results import FlowOUTrackingResult
    trk = object.__new__(M.MFT)
    C = Config(); C.deltas = deltas; C.occlusion_threshold = 0.02
    trk.C, trk.device = C, 'cpu'
    eng = _FakeEngine()
    trk.flower = type('F', (), {'ensure_geometry': lambda self, H, W, claim=False: eng})()
    calls = []

    def fake_refine(lefts, rights, out=None):
        calls.append((list(lefts), list(rights)))
        t = out if out is not None else torch.zeros((len(lefts), 4, 8, 8))
        t.zero_()
        return t
    eng.refine = fake_refine
    monkeypatch.setattr(M, 'chain_select', lambda lefts, right, thr, want_index=False: (torch.zeros((4, 8, 8)), None))
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda: type('S', (), {'synchronize': lambda self: None})())
    return trk, eng, calls


@pytest.mark.parametrize('direction', [1, -1])
def test_tracker_bookkeeping(direction, monkeypatch):
    """Live chains, dedup, memory eviction and feature-slot accounting over 80 frames
    (MFT.py:74-91,157-185) against the oracle's bookkeeping."""
    deltas = [np.inf, 1, 2, 4, 8, 16, 32]
    start = 100
    trk, eng, calls = _tracker_without_gpu(deltas, direction, start, monkeypatch)
    img = lambda i: np.full((8, 8, 3), i % 251, np.uint8)
    trk.init(img(start), start_frame_i=start, time_direction=direction)
    for n in range(1, 80):
        t = start + n * direction
        meta = trk.track(img(t))
        want = O.live_chains(deltas, t, start, direction)
        assert trk.live_chains() == want
        lefts, rights = calls[-1]
        assert len(lefts) == len(want) and len(set(rights)) == 1
        # every pair's slots hold the frames they should
        assert [eng.encoded[s] for s in lefts] == [left % 251 for _, left in want]
        assert eng.encoded[rights[0]] == t % 251
        keep = {start} | {f for f in range(min(start, t), max(start, t) + 1) if abs(t - f) < 32}
        assert set(trk.memory) == keep
        slots = [m['slot'] for m in trk.memory.values()]
        assert len(set(slots)) == len(slots) and not (set(slots) & set(trk._free_slots))
        assert tuple(meta.result.flow.shape) == (2, 8, 8)
        assert len(trk._pool.free) >= 4                 # dropped results hand their pinned slot back
    assert [len(c[0]) for c in calls[:5]] == [1, 2, 3, 3, 4] and len(calls[40][0]) == 7


def test_results_cpu_paths_match_oracle():
    from mft_b200.results import FlowOUTrackingResult
    rng = np.random.default_rng(0)
    H, W = 20, 30
    l = np.concatenate([rng.standard_normal((2, H, W)) * 3, rng.uniform(0, 1, (2, H, W))]).astype(np.float32)
    r = np.concatenate([rng.standard_normal((2, H, W)) * 3, rng.uniform(0, 1, (2, H, W))]).astype(np.float32)
    res = FlowOUTrackingResult.from_packed(torch.from_numpy(l.copy()))
    wf, _, _ = O.chain((l[:2], l[2:3], l[3:4]), (r[:2], r[2:3], r[3:4]))
    assert np.abs(res.chain(torch.from_numpy(r[:2])).numpy() - wf).max() < 1e-4
    assert res.packed().data_ptr() == res.flow.data_ptr()            # views, no copy
    c = res.clone(); c.occlusion[:] = 1
    assert float(res.occlusion.max()) < 1
    ident = FlowOUTrackingResult.identity((H, W))
    assert float(ident.packed().abs().max()) == 0 and not ident.invalid_mask().any()
    pts = torch.tensor([[0.0, 0.0], [W - 1.0, H - 1.0], [3.5, 4.25]])
    assert torch.allclose(ident.warp_forward_points(pts), pts)
    img = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    # zero flow splats in place -- except the last row / column, which the reference's clamped corner indices give zero
    # weight (interpolation.py:256-278: x0 == x1 == W-1 there)
    wfz = ident.warp_forward(img)
    assert np.abs(wfz[:-1, :-1] - img[:-1, :-1]).max() < 1e-6 and np.abs(wfz[-1]).max() == 0 and np.abs(wfz[:, -1]).max() == 0
    assert np.abs(wfz - O.warp_forward(np.zeros((2, H, W), np.float32), img)).max() < 1e-6


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from mft_b200.dist import FlowShardedTracker
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    out.put((rank, (_run_sharded(world), _run_delta_sharded(world))))
    dist.destroy_process_group()


def _run_delta_sharded(world):
    """Per-delta sharding inside a frame: stand-in flow / select functions that depend on the chain
    identity and on the ORDER of the gathered fields, so a wrong owner, slot or order changes the sums."""
    from mft_b200.dist import DeltaShardedTracker
    H, W, T = 5, 7, 40
    deltas = [np.inf, 1, 2, 4, 8, 16, 32]
    calls = []

    def flow_fn(t, live, out=None):
        calls.append(len(live))
        return torch.stack([torch.full((4, H, W), float(t * 100 + (0 if np.isinf(d) else d)) + 0.25 * left) for d, left in live])

    def select_fn(lefts, right):
        w = torch.arange(1, len(lefts) + 1, dtype=torch.float32).view(-1, 1, 1, 1)
        return (torch.stack(lefts) * 0.125 + right * w).sum(0) / float(len(lefts)) % 977.0
    trk = DeltaShardedTracker(deltas, (H, W), flow_fn, select_fn, 'cpu')
    sums = [float(trk.track().sum()) for _ in range(T)]
    assert max(calls) <= (7 + world - 1) // world              # no rank refines more than ceil(K / G) pairs per frame
    assert 0 in trk.results and len(trk.results) <= 33 and min(k for k in trk.results if k) == T - 31      # template + last 32
    return sums


def _run_sharded(world):
    """Per-timestep flow sharding as bench.py --mode flow-shard drives it: the encoders are sharded too (rank t % G "encodes"
    frame t, one feature all_gather per round through encode_fn), the flows depend on the GATHERED features of both frames,
    and the scan continues across two run_range calls (state-building part, then the timed part)."""
    import torch.distributed as dist
    from mft_b200.dist import FlowShardedTracker
    H, W, T = 6, 8, 13
    deltas = [np.inf, 1, 2, 4]
    rank = dist.get_rank() if dist.is_initialized() else 0
    feats = torch.zeros(T + world, 3)
    feats[0] = 0.5                                             # the template: every rank encodes it itself
    encoded = []

    def encode_fn(ts):
        t0 = ts[0]
        mine = t0 + rank
        if mine < T:
            feats[mine] = 1.5 * mine + 0.25                    # only the owner computes the frame's features
            encoded.append(mine)
        if world > 1:
            blk = feats[t0:t0 + world]
            dist.all_gather_into_tensor(blk.view(world * 3), feats[t0 + rank].clone())

    def flow_fn(t, live, out=None):        # deterministic stand-in for the batched refinement: reads both frames' features
        return torch.stack([torch.full((4, H, W), float(t * 10 + (0 if np.isinf(d) else d)) + float(feats[left].sum() * 0.125 + feats[t].sum()))
                            for d, left in live])

    def select_fn(lefts, right):  # deterministic stand-in for chain_select
        return sum(l * 0.5 for l in lefts) / len(lefts) + right.mean(0)
    trk = FlowShardedTracker(deltas, T, (H, W), flow_fn, select_fn, 'cpu', encode_fn=encode_fn)
    split = 1 + 3 * world                                      # a round boundary
    trk.run_range(1, split)
    res = trk.run_range(split, T)
    assert encoded == [t for t in range(1, T) if (t - 1) % world == rank]      # every frame encoded once, by its owner
    return {k: float(v.sum()) for k, v in res.items()}


def test_flow_sharding_world2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    want = (_run_sharded(1), _run_delta_sharded(1))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == want and got[1] == want
    from mft_b200.dist import shard_items
    assert shard_items(10, 1, 4) == [1, 5, 9] and sorted(sum((shard_items(10, r, 4) for r in range(4)), [])) == list(range(10))


def test_flow_cache_protocol_host_logic(monkeypatch):
    """Cached pairs are not recomputed; only finite deltas are cached unless C.cache_delta_infinity (MFT.py:99,209-228)."""
    deltas = [np.inf, 1, 2]
    trk, eng, calls = _tracker_without_gpu(deltas, 1, 0, monkeypatch)

    class Cache:
        def __init__(self):
            self.store, self.writes = {}, []

        def read(self, l, r):
            return self.store.get((l, r), (None, None, None))

        def write(self, l, r, f, o, s):
            self.writes.append((l, r))
            self.store[(l, r)] = (f, o, s)
    cache = Cache()
    img = lambda i: np.full((8, 8, 3), i, np.uint8)
    for rep in range(2):
        trk.init(img(0), flow_cache=cache)
        n0 = len(calls)
        for t in range(1, 5):
            trk.track(img(t))
        pairs = [len(c[0]) for c in calls[n0:]]
        # first pass computes every live chain; second pass only the (uncached) template chain
        assert pairs == ([1, 2, 3, 3] if rep == 0 else [1, 1, 1, 1])
    # (0,1) is computed under delta=inf (delta 1 collapses onto it, MFT.py:90) and is therefore not cached
    assert sorted(set(cache.writes)) == [(1, 2), (1, 3), (2, 3), (2, 4), (3, 4)]
    trk.C.cache_delta_infinity = True
    trk.init(img(0), flow_cache=cache)
    trk.track(img(1))
    assert (0, 1) in cache.store


def test_warp_forward_cpu_path_matches_reference_golden():
    """FlowOUTrackingResult.warp_forward on a CPU result (what demo.py holds) follows the reference's clamped splat."""
    import numpy as np
    import torch
    from conftest import golden
    from mft_b200.results import FlowOUTrackingResult
    g = golden('warp_forward.npz')
    r = FlowOUTrackingResult(torch.from_numpy(g['flow']))
    assert np.abs(r.warp_forward(g['img']) - g['out_plain']).max() < 1e-6
    assert np.abs(r.warp_forward(g['img'], mask=g['mask'], border=-1.0) - g['out_mask']).max() < 1e-6


def test_device_flow_cache_protocol_on_cpu_tensors():
    """read / write protocol of the reference FlowCache (io.py:655-698), LRU eviction under a byte budget, fp16 storage."""
    import torch
    from mft_b200.flow_cache import DeviceFlowCache
    H, W = 8, 12
    one = 4 * H * W * 4
    c = DeviceFlowCache(max_bytes=3 * one, device='cpu')
    assert c.read(0, 1) == (None, None, None)
    for i in range(5):
        c.write(i, i + 1, torch.full((2, H, W), float(i)), torch.zeros(1, H, W), torch.ones(1, H, W))
    assert len(c) == 3 and c.evictions == 2 and c.read(0, 1)[0] is None
    f, o, s = c.read(3, 4)
    assert tuple(f.shape) == (2, H, W) and tuple(o.shape) == (1, H, W) and tuple(s.shape) == (1, H, W) and f.mean() == 3
    c.write(9, 10, torch.zeros(2, H, W), torch.zeros(1, H, W), torch.ones(1, H, W))      # evicts (2,3): (3,4) was just used
    assert c.read(2, 3)[0] is None and c.read(3, 4)[0] is not None
    h = DeviceFlowCache(dtype=torch.float16, device='cpu')
    h.write(0, 1, torch.full((2, H, W), 1.5), torch.zeros(1, H, W), torch.ones(1, H, W))
    assert h.bytes == one // 2 and h.read(0, 1)[0].dtype == torch.float32 and h.read(0, 1)[0].mean() == 1.5


def test_point_queries_cpu_paths_match_reference_golden():
    """Result geometry + convert_to_point_tracking on CPU tensors (what callers hold after track()) against vectors recorded
    from the unmodified reference (results.py:87-188,250-265; point_tracking.py:6-27): queries inside, on the border, outside."""
    from conftest import golden
    from mft_b200.point_tracking import convert_to_point_tracking
    from mft_b200.results import FlowOUTrackingResult
    g = golden('point_queries.npz')
    res = FlowOUTrackingResult(torch.from_numpy(g['flow']), torch.from_numpy(g['occlusion']), torch.from_numpy(g['sigma']))
    q = torch.from_numpy(g['queries'])
    assert np.abs(res.warp_forward_points(q).numpy() - g['warped_points']).max() < 1e-4
    f, o, s = res.sample(q)
    assert tuple(f.shape) == g['sample_flow'].shape and tuple(o.shape) == g['sample_occlusion'].shape
    assert np.abs(f.numpy() - g['sample_flow']).max() < 1e-4
    assert np.abs(o.numpy() - g['sample_occlusion']).max() < 1e-5 and np.abs(s.numpy() - g['sample_sigma']).max() < 1e-5
    pc, po = convert_to_point_tracking(res, g['queries'])
    assert pc.shape == g['pt_coords'].shape and po.shape == g['pt_occlusion'].shape and po.dtype == np.float32
    assert np.abs(pc - g['pt_coords']).max() < 1e-4 and np.abs(po - g['pt_occlusion']).max() < 1e-5
    assert np.abs(res.chain(torch.from_numpy(g['other'])).numpy() - g['chained']).max() < 1e-4
    assert np.abs(res.warp_backward(torch.from_numpy(g['img'])).numpy() - g['warped_img']).max() < 1e-5
    assert np.array_equal(res.invalid_mask().numpy(), g['invalid'])


def test_roofline_flop_counts_follow_the_layer_shapes():
    """bench.py's algorithmic FLOPs (roofline numerators) recomputed from the checkpoint's layer shapes."""
    import bench
    spec = {n[:-7]: s for n, s in O.weight_spec() if n.endswith('.weight') and len(s) == 4}
    f = lambda name: 2 * int(np.prod(spec[name]))                         # 2 * cout * cin * kh * kw per output pixel
    ub = 'update_block.'
    per_iter = sum(f(ub + n) for n in ('encoder.convc1', 'encoder.convc2', 'encoder.convf1', 'encoder.convf2', 'encoder.conv',
                                       'gru.convz1', 'gru.convr1', 'gru.convq1', 'gru.convz2', 'gru.convr2', 'gru.convq2',
                                       'flow_head.conv1', 'flow_head.conv2'))
    assert per_iter == 5351936                                             # one conv_prog_kernel launch, per coarse pixel
    last = per_iter + f(ub + 'mask.0') + f(ub + 'mask.2')
    ou = sum(f('occlusion_block.' + n) for n in ('occl_head.conv1', 'occl_head.conv2', 'uncertainty_head.conv1', 'uncertainty_head.conv2'))
    assert (last, ou) == (6236672, 3287808)
    enc2 = f('fnet.conv1') + sum(f(f'fnet.layer1.{b}.conv{c}') for b in (0, 1) for c in (1, 2))
    assert enc2 == 18816 + 4 * 73728
    H = W = 512
    n = (H // 8) * (W // 8)
    F = bench.flops_per_frame(H, W)
    assert abs(F - 2.09e12) < 0.01e12
    assert F == 2 * ((H // 2) * (W // 2) * enc2 + (H // 4) * (W // 4) * 620544 + n * (1130496 + 65536)) + \
        7 * (2 * n * n * 256 + 11 * n * per_iter + n * last + n * ou)


def test_tapvid_runner_host_logic():
    """mft_b200.tapvid mirrors run_MFT_tapvid.py:140-285: per start frame a forward (+ backward in 'strided' mode) init/track
    sequence, every frame's result sampled at that start frame's queries.  Exercised with a stand-in tracker (no GPU)."""
    from types import SimpleNamespace
    from mft_b200 import tapvid as TV
    from mft_b200.results import FlowOUTrackingResult

    class FakeTracker:
        def __init__(self):
            self.log = []

        def init(self, img, start_frame_i=0, time_direction=1, flow_cache=None):
            self.t, self.dir, self.start = start_frame_i, time_direction, start_frame_i
            self.log.append(('init', start_frame_i, time_direction, int(img[0, 0, 0])))
            return SimpleNamespace(result=FlowOUTrackingResult.identity(img.shape[:2]))

        def track(self, img, debug=False, device_result=False):
            self.t += self.dir
            self.log.append(('track', self.t, int(img[0, 0, 0])))
            r = FlowOUTrackingResult.identity(img.shape[:2])
            r.flow[0] += float(self.t - self.start)           # x flow = signed frame distance from the start frame
            return SimpleNamespace(result=r)

    data = TV.synthetic_dataset(1, 12, 6, 32, seed=3)
    d = data['synth-000']
    assert d['video'].shape == (12, 32, 32, 3) and d['video'].dtype == np.uint8
    assert d['points'].shape == (6, 12, 2) and d['occluded'].shape == (6, 12) and d['occluded'].dtype == np.bool_
    assert float(d['points'].min()) >= 0 and float(d['points'].max()) <= 1
    video = np.stack([np.full((32, 32, 3), i, np.uint8) for i in range(12)])          # frame i is filled with i
    pts = d['points'] * 32
    qf, qs = TV.sample_queries_first(d['occluded'], pts), TV.sample_queries_strided(d['occluded'], pts)
    assert qf.shape[1] == 3 and all(not d['occluded'][i, int(q[0])] for i, q in enumerate(qf))
    assert set(np.unique(qs[:, 0]).astype(int)) <= {0, 5, 10}
    trk = FakeTracker()
    tracks, occl, n = TV.run_sequence(trk, video, qs, 'strided', device='cpu')
    starts = sorted(set(qs[:, 0].astype(int)))
    assert n == sum((12 - s) + (s + 1) for s in starts)
    inits = [e for e in trk.log if e[0] == 'init']
    assert [(e[1], e[2]) for e in inits] == [(s, dirn) for s in starts for dirn in (1, -1)]
    assert all(e[3] == e[1] for e in inits) and all(e[2] == e[1] for e in trk.log if e[0] == 'track')     # frame i went to time i
    # a query started at frame s sits at x + (t - s) in frame t (the stand-in's flow), on both sides of s
    q = qs.astype(np.int64)
    for k in range(len(q)):
        s, y, x = q[k]
        for t in range(12):
            assert abs(tracks[k, t, 0] - (x + (t - s))) < 1e-4 and abs(tracks[k, t, 1] - y) < 1e-4
    tracks_f, _, n_f = TV.run_sequence(FakeTracker(), video, qf, 'first', device='cpu')
    assert n_f == sum(12 - s for s in sorted(set(qf[:, 0].astype(int))))


def test_demo_video_helper():
    """bench.py --mode demo (BASELINE config 2) reads the reference's demo video as DATA through mft_b200.synth."""
    from mft_b200.synth import demo_video_frames, find_demo_video
    if find_demo_video() is None:
        pytest.skip('demo video not reachable here')
    fr = demo_video_frames((96, 64), max_frames=3)
    assert len(fr) == 3 and fr[0].shape == (64, 96, 3) and fr[0].dtype == np.uint8 and fr[0].flags.c_contiguous
    assert any(not np.array_equal(fr[0], f) for f in fr[1:])


def test_bench_modes_parse():
    """The bench's command line keeps the driver's contract (defaults: one GPU, 20 steps, 3 warm-up) and names every mode."""
    import bench
    assert bench.parse_size('', 512) == (512, 512) and bench.parse_size('1080x1920', 512) == (1080, 1920) and bench.parse_size('1024', 512) == (1024, 1024)
    src = open(bench.__file__).read()
    for mode in ('track', 'flow-shard', 'tapvid', 'demo'):
        assert f"'{mode}'" in src
    assert abs(bench.flops_per_frame(512, 512) - 2092426067968) < 1
