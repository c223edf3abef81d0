"""First-contact GPU probe: runs each kernel family against a torch / oracle reference and prints
diagnostics without stopping at the first failure.  Usage: python tools/gpu_probe.py [stage ...]"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mft_b200 import engine as E, weights as WT   # noqa: E402
from oracle import mft_oracle as O                # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def conv_case(cin, cout, kh, kw, stride, H, W, B, n_tile, impl, pitch=None, seed=0, verbose=True):
    g = torch.Generator().manual_seed(seed)
    pitch = pitch or (cin + 7) // 8 * 8
    x = torch.randn(B, H, W, pitch, generator=g).half()
    w = (torch.randn(cout, cin, kh, kw, generator=g) / np.sqrt(cin * kh * kw)).half().float()
    b = torch.randn(cout, generator=g)
    w16, bias, cout_pad, ktot, bias_len = WT._pack(w, b, cout_pad=(cout + n_tile - 1) // n_tile * n_tile)
    xd = x.cuda()
    wd = torch.from_numpy(w16.view(np.float16)).cuda()
    bd = torch.from_numpy(bias).cuda()
    out = E.conv2d_test(xd, wd, bd, cin, cout_pad, n_tile, kh, kw, stride, False, impl)
    ref = torch.nn.functional.conv2d(xd[..., :cin].float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), stride=stride,
                                     padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
    got = out[..., :cout]
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err < 2e-3 * max(scale, 1)
    if verbose or not ok:
        print(f'conv impl={impl} cin={cin} cout={cout} k={kh}x{kw} s={stride} HxW={H}x{W} B={B} n_tile={n_tile}: '
              f'maxerr={err:.3e} scale={scale:.2f} {"OK" if ok else "FAIL"}', flush=True)
        if not ok:
            bad = ((got - ref).abs() > 2e-3 * max(scale, 1))
            idx = bad.nonzero()
            print('   bad count', int(bad.sum()), 'of', bad.numel(), 'first', idx[:5].tolist(), flush=True)
            print('   got', got[tuple(idx[0])].item(), 'ref', ref[tuple(idx[0])].item())
            print('   per-channel bad', bad.sum((0, 1, 2))[:16].tolist(), 'per-row bad', bad.sum((0, 2, 3))[:16].tolist())
    return ok


def stage_conv_simt():
    ok = True
    for cfg in [(64, 64, 1, 1, 1, 8, 16, 1, 64), (64, 64, 3, 3, 1, 16, 16, 1, 64), (324, 256, 1, 1, 1, 16, 16, 2, 256),
                (64, 96, 3, 3, 2, 32, 32, 1, 96), (384, 256, 1, 5, 1, 16, 24, 1, 256), (128, 2, 3, 3, 1, 16, 16, 1, 16)]:
        ok &= conv_case(*cfg, impl=1)
    return ok


def stage_conv_tc():
    ok = True
    cfgs = [(64, 64, 1, 1, 1, 8, 16, 1, 64),        # single tile, single stage, no taps
            (128, 64, 1, 1, 1, 8, 16, 1, 64),       # 2 k-chunks
            (64, 64, 3, 3, 1, 8, 16, 1, 64),        # taps / OOB zero fill
            (64, 64, 3, 3, 1, 16, 16, 1, 64),       # 2 tiles
            (64, 128, 3, 3, 1, 24, 40, 2, 128),     # ragged tiles, batch
            (324, 256, 1, 1, 1, 16, 16, 2, 256),    # cin not multiple of 64
            (256, 192, 3, 3, 1, 16, 16, 1, 192),
            (384, 256, 1, 5, 1, 16, 24, 1, 256),
            (384, 128, 5, 1, 1, 16, 24, 1, 128),
            (256, 2, 3, 3, 1, 16, 16, 1, 16),
            (256, 576, 1, 1, 1, 16, 16, 1, 192),    # 3 n-tiles
            (147, 64, 1, 1, 1, 1, 1024, 1, 64),     # flat GEMM view
            (712, 256, 3, 3, 1, 16, 16, 1, 256),
            (64, 96, 3, 3, 2, 32, 32, 1, 96),       # stride 2 via TMA element strides
            (64, 96, 1, 1, 2, 32, 32, 1, 96),
            (96, 128, 3, 3, 2, 64, 64, 1, 128)]
    for cfg in cfgs:
        try:
            ok &= conv_case(*cfg, impl=0)
        except Exception as ex:
            ok = False
            print('conv tc', cfg, 'EXC', ex, flush=True)
    return ok


def stage_chain_select():
    ok = True
    rng = np.random.default_rng(7)
    for (H, W, K) in ((24, 40, 7), (130, 258, 3), (512, 512, 7)):
        lefts = [np.concatenate([(rng.standard_normal((2, H, W)) * 5), rng.uniform(0, 0.04, (1, H, W)),
                                 rng.uniform(0, 2, (1, H, W))]).astype(np.float32) for _ in range(K)]
        right = np.stack([np.concatenate([(rng.standard_normal((2, H, W)) * 3), rng.uniform(0, 0.03, (1, H, W)),
                                          rng.uniform(0.05, 2, (1, H, W))]).astype(np.float32) for _ in range(K)])
        lefts[0][2, :3] = 0.5
        for k in range(K):
            lefts[k][:2, 5:8] = np.round(lefts[k][:2, 5:8])
        lefts[0][0, 8:10] = 1000
        t = time.time()
        cands = [O.chain((l[:2], l[2:3], l[3:4]), (r[:2], r[2:3], r[3:4])) for l, r in zip(lefts, right)]
        f, o, s, idx = O.select(cands, 0.02)
        want = np.concatenate([f, o, s])
        out, gi = E.chain_select([torch.from_numpy(l).cuda() for l in lefts], torch.from_numpy(right).cuda(), 0.02)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        same = (got == want) | (np.isnan(got) & np.isnan(want))
        iok = (gi.cpu().numpy() == idx).all()
        print(f'chain_select {H}x{W} K={K}: bit-exact values {same.mean():.6f} index exact {iok} '
              f'maxabs {np.nanmax(np.abs(got - want)):.3e}', flush=True)
        ok &= bool(same.all() and iok)
    return ok


def stage_raft(tag='seeded', size=128):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', f'raft_{tag}_128.npz'))
    frames = dict(zip(g['frame_ids'].tolist(), g['frames']))
    if tag == 'seeded':
        W = O.seeded_weights(0)
    else:
        from oracle import fetch_ref_assets
        W = O.load_checkpoint(fetch_ref_assets.find_checkpoint())
    eng = E.Engine(W)
    eng.configure(128, 128, max_pairs=2, n_slots=4, iters=12)
    impl = int(os.environ.get('PROBE_CONV_IMPL', '0'))
    eng.set_option('conv_impl', impl)
    for slot, fid in enumerate((0, 1, 8)):
        eng.encode_frame(frames[fid], slot)
    eng.check_device()
    taps = {}
    of, oo, os_ = O.compute_flow(W, frames[0], frames[1], taps=taps)
    npx = 256
    fm = eng.debug_buffer('fmap_slots', torch.float16, (4, npx, 256)).float().cpu()
    ref_f1 = taps['fmap1'][0].reshape(256, npx).t()
    print('fmap slot0 err', (fm[0] - ref_f1).abs().max().item(), 'scale', ref_f1.abs().max().item())
    ref_f2 = taps['fmap2'][0].reshape(256, npx).t()
    print('fmap slot1 err', (fm[1] - ref_f2).abs().max().item())
    net = eng.debug_buffer('net_slots', torch.float32, (4, npx, 128)).cpu()
    print('net0 err', (net[0] - taps['net0'][0].reshape(128, npx).t()).abs().max().item())
    inp = eng.debug_buffer('inp_slots', torch.float16, (4, npx, 128)).float().cpu()
    print('inp err', (inp[0] - taps['inp'][0].reshape(128, npx).t()).abs().max().item(), 'scale', taps['inp'].abs().max().item())
    for it in (1, 2, 12):
        eng.set_option('iters', it)
        out = eng.refine([0, 0], [1, 2])
        eng.check_device()
        c1 = eng.debug_buffer('coords1', torch.float32, (2, npx, 2)).cpu()
        ref_c = taps['iters'][it - 1]['coords1'].reshape(2, npx).t()
        print(f'iters={it}: coords1 err', (c1[0] - ref_c).abs().max().item())
        if it == 1:
            l0 = eng.debug_buffer('corr_l0', torch.float32, (npx, npx)).cpu()
            print('   corr l0 err', (l0 - taps['pyramid'][0].reshape(npx, npx)).abs().max().item(), 'scale', taps['pyramid'][0].abs().max().item())
            l3 = eng.debug_buffer('corr_l3', torch.float32, (npx, 4)).cpu()
            print('   corr l3 err', (l3 - taps['pyramid'][3].reshape(npx, 4)).abs().max().item())
            c16 = eng.debug_buffer('corr16', torch.float16, (npx, 328)).float().cpu()
            ref_corr = taps['iters'][0]['corr'][0].reshape(324, npx).t()
            print('   lookup err', (c16[:, :324] - ref_corr).abs().max().item())
            h32 = eng.debug_buffer('h32', torch.float32, (npx, 128)).cpu()
            print('   net err', (h32 - taps['iters'][0]['net'][0].reshape(128, npx).t()).abs().max().item())
            X = eng.debug_buffer('X', torch.float16, (npx, 512)).float().cpu()
            ref_m = taps['iters'][0]['motion'][0].reshape(128, npx).t()
            print('   motion err', (X[:, 256:384] - ref_m).abs().max().item(), 'scale', ref_m.abs().max().item())
    o = out.cpu()
    print('flow err (0,1)', (o[0, :2] - of).abs().max().item(), 'mean', (o[0, :2] - of).abs().mean().item(),
          'occ err', (o[0, 2:3] - oo).abs().max().item(), 'sigma relerr', ((o[0, 3:4] - os_).abs() / (1 + os_)).max().item())
    print('vs golden flow (0,8)', np.abs(o[1, :2].numpy() - g['flow_0_8']).max(), 'mean', np.abs(o[1, :2].numpy() - g['flow_0_8']).mean())
    print('launches', eng.launch_count())
    return True


STAGES = {'conv_simt': stage_conv_simt, 'conv_tc': stage_conv_tc, 'chain_select': stage_chain_select,
          'raft_seeded': lambda: stage_raft('seeded'), 'raft_real': lambda: stage_raft('real')}

if __name__ == '__main__':
    names = sys.argv[1:] or list(STAGES)
    print(torch.cuda.get_device_name(0), 'cpus', os.cpu_count(), flush=True)
    for n in names:
        print(f'===== {n}', flush=True)
        t = time.time()
        try:
            r = STAGES[n]()
        except Exception:
            traceback.print_exc()
            r = False
        print(f'===== {n}: {"PASS" if r else "FAIL"} ({time.time() - t:.1f}s)', flush=True)
