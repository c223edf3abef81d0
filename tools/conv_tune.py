"""Conv-kernel tuning sweep on the hot layer shapes of the 512x512, 7-pair refinement loop."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mft_b200 import engine as E, weights as WT   # noqa: E402

torch.backends.cudnn.allow_tf32 = False

LAYERS = {   # name: (cin, cout, kh, kw, n_tile)
    'convc1': (324, 256, 1, 1, 256), 'convc2': (256, 192, 3, 3, 192), 'convf1': (98, 128, 1, 1, 128),
    'convf2': (128, 64, 3, 3, 64), 'convm': (256, 126, 3, 3, 128), 'gru_zr': (384, 256, 1, 5, 256),
    'gru_q': (384, 128, 5, 1, 128), 'fh1': (128, 256, 3, 3, 256), 'fh2': (256, 2, 3, 3, 16), 'ou1': (712, 256, 3, 3, 256), 'corr': (256, 4096, 1, 1, 256),
}


def run(name, cluster, smem, B=7, H=64, W=64, reps=20, check=True):
    cin, cout, kh, kw, n_tile = LAYERS[name]
    g = torch.Generator().manual_seed(1)
    pitch = (cin + 7) // 8 * 8
    x = torch.randn(B, H, W, pitch, generator=g).half().cuda()
    w = (torch.randn(cout, cin, kh, kw, generator=g) / np.sqrt(cin * kh * kw)).half().float()
    b = torch.randn(cout, generator=g)
    w16, bias, cout_pad, ktot, _ = WT._pack(w, b, cout_pad=(cout + n_tile - 1) // n_tile * n_tile)
    wd = torch.from_numpy(w16.view(np.float16)).cuda()
    bd = torch.from_numpy(bias).cuda()
    timing = torch.zeros((4096, 16), dtype=torch.int64, device='cuda') if os.environ.get('TUNE_TIMING') else None
    try:
        out, ms = E.conv2d_bench(x, wd, bd, cin, cout_pad, n_tile, kh, kw, 1, False, cluster, smem, reps, timing)
    except Exception as ex:
        return f'{name:7s} cluster={cluster} smem={smem}: EXC {ex}'
    extra = ''
    if timing is not None:
        t = timing.cpu().numpy()
        t = t[t[:, 0] != 0].astype(np.float64)
        d = lambda a, b: (t[:, a] - t[:, b]).mean()
        span = (t[:, 7].max() - t[:, 7].min()) / 1e3
        extra = (f'\n        ctas={len(t)} cycles: setup {d(1,0):.0f} fill {d(2,1):.0f} issue {d(3,2):.0f} '
                 f'accum_ready-after-first {d(4,2):.0f} epilogue {d(5,4):.0f} teardown {d(6,5):.0f} total {d(6,0):.0f}; '
                 f'CTA start spread {span:.1f} us'
                 f'\n        chunk0: tmem_ld {d(9,8):.0f} stage {d(10,9):.0f} store {d(11,10):.0f} | chunk1: gap {d(12,11):.0f} tmem_ld {d(13,12):.0f} stage {d(14,13):.0f} store {d(15,14):.0f}')
    err = -1.0
    if check:
        ref = torch.nn.functional.conv2d(x[..., :cin].float().permute(0, 3, 1, 2), w.cuda(), b.cuda(),
                                         padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
        err = (out[..., :cout] - ref).abs().max().item()
    flops = 2.0 * B * H * W * cin * kh * kw * cout
    return f'{name:7s} cluster={cluster} smem={smem:3d}: {ms * 1e3:7.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s  maxerr={err:.2e}' + extra


if __name__ == '__main__':
    if os.environ.get('TUNE_V2') is not None:
        E.set_global_option('conv_v2', int(os.environ['TUNE_V2']))
    names = sys.argv[1].split(',') if len(sys.argv) > 1 else list(LAYERS)
    clusters = [int(c) for c in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['1', '2', '4', '8'])]
    smems = [int(c) for c in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['200'])]
    for n in names:
        for c in clusters:
            for s in smems:
                print(run(n, c, s), flush=True)
