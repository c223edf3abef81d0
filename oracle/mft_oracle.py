"""CPU oracle for the MFT per-frame tracking hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement of the algorithm of serycjon/MFT's hot
path (RAFT-OU forward per chain delta, flow-chain composition, per-pixel best-chain
selection).  It exists so that the CUDA path in ``mft_b200`` has something to be
checked against on a box where ``/root/reference`` does not exist.

* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
  may import this module.  The product package ``mft_b200`` never does.
* Parity pin: the reference ships NO tests / golden vectors, so the
  oracle is pinned against outputs of the reference itself, run on CPU in the build
  container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``; checked by
  ``tests/test_oracle_golden.py``).
* Arithmetic: convolutions / matmul use torch fp32 CPU ops (the reference bottoms
  out in the same ATen calls); every gather / bilinear / selection
  step is restated explicitly with a *defined fp32 operation order* (numpy float32,
  no fused multiply-add) which the CUDA kernels reproduce bit for bit.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32

# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------

#: (name, shape) of every tensor the hot path consumes, in checkpoint order
#: (checkpoints/*.pth, keys 'module.<name>').
def weight_spec():
    spec = []

    def conv(name, co, ci, kh, kw):
        spec.append((name + '.weight', (co, ci, kh, kw)))
        spec.append((name + '.bias', (co,)))

    def bn(name, c):
        for s in ('weight', 'bias', 'running_mean', 'running_var'):
            spec.append((f'{name}.{s}', (c,)))

    for net in ('fnet', 'cnet'):
        has_bn = net == 'cnet'
        if has_bn:
            bn(f'{net}.norm1', 64)
        conv(f'{net}.conv1', 64, 3, 7, 7)
        cin = 64
        for li, dim in ((1, 64), (2, 96), (3, 128)):
            for bi in (0, 1):
                p = f'{net}.layer{li}.{bi}'
                conv(p + '.conv1', dim, cin if bi == 0 else dim, 3, 3)
                conv(p + '.conv2', dim, dim, 3, 3)
                if has_bn:
                    bn(p + '.norm1', dim)
                    bn(p + '.norm2', dim)
                if bi == 0 and li > 1:
                    if has_bn:
                        bn(p + '.norm3', dim)
                    conv(p + '.downsample.0', dim, cin, 1, 1)
            cin = dim
        conv(f'{net}.conv2', 256, 128, 1, 1)
    ub = 'update_block'
    conv(f'{ub}.encoder.convc1', 256, 324, 1, 1)
    conv(f'{ub}.encoder.convc2', 192, 256, 3, 3)
    conv(f'{ub}.encoder.convf1', 128, 2, 7, 7)
    conv(f'{ub}.encoder.convf2', 64, 128, 3, 3)
    conv(f'{ub}.encoder.conv', 126, 256, 3, 3)
    for g in 'zrq':
        conv(f'{ub}.gru.conv{g}1', 128, 384, 1, 5)
    for g in 'zrq':
        conv(f'{ub}.gru.conv{g}2', 128, 384, 5, 1)
    conv(f'{ub}.flow_head.conv1', 256, 128, 3, 3)
    conv(f'{ub}.flow_head.conv2', 2, 256, 3, 3)
    conv(f'{ub}.mask.0', 256, 128, 3, 3)
    conv(f'{ub}.mask.2', 576, 256, 1, 1)
    ob = 'occlusion_block'
    conv(f'{ob}.occl_head.conv1', 128, 712, 3, 3)
    conv(f'{ob}.occl_head.conv2', 2, 128, 3, 3)
    conv(f'{ob}.uncertainty_head.conv1', 128, 712, 3, 3)
    conv(f'{ob}.uncertainty_head.conv2', 1, 128, 3, 3)
    return spec


def seeded_weights(seed=0):
    """Deterministic stand-in weights with the checkpoint's exact shapes.

    Used where the real checkpoint cannot travel (GPU box).  He-style fan-in scaling keeps
    activations O(1) through the network so the numerics are exercised realistically; the
    GRU / head convs are scaled down a little so 12 iterations stay bounded."""
    rng = np.random.default_rng(seed)
    W = {}
    for name, shape in weight_spec():
        if name.endswith('.weight') and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            gain = math.sqrt(2.0 / fan_in)
            if '.gru.' in name:
                gain *= 0.7
            if 'flow_head.conv2' in name or 'head.conv2' in name:
                gain *= 0.25
            W[name] = torch.from_numpy((rng.standard_normal(shape) * gain).astype(F32))
        elif name.endswith('running_var'):
            W[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(F32))
        elif name.endswith('running_mean'):
            W[name] = torch.from_numpy((rng.standard_normal(shape) * 0.1).astype(F32))
        elif name.endswith('.weight'):          # BN gamma
            W[name] = torch.from_numpy(rng.uniform(0.8, 1.2, shape).astype(F32))
        else:                                   # conv bias / BN beta
            W[name] = torch.from_numpy((rng.standard_normal(shape) * 0.05).astype(F32))
    return W


def load_checkpoint(path):
    """Checkpoint loader: strips the DataParallel 'module.' prefix (MFT/raft.py:20-23)."""
    sd = torch.load(path, map_location='cpu', weights_only=True)
    W = {}
    for k, v in sd.items():
        if k.startswith('module.'):
            k = k[len('module.'):]
        if k.endswith('num_batches_tracked'):
            continue
        W[k] = v.float().contiguous()
    return W


# ----------------------------------------------------------------------------------------------
# encoders (MFT/RAFT/core/extractor.py:6-56, 118-195)
# ----------------------------------------------------------------------------------------------

def _conv(x, W, name, stride=1, padding=0):
    return F.conv2d(x, W[name + '.weight'], W[name + '.bias'], stride=stride, padding=padding)


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm2d defaults: per image, per channel, biased variance, no affine
    (extractor.py:28-32,129-130)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def batch_norm_eval(x, W, name, eps=1e-5):
    g, b = W[name + '.weight'], W[name + '.bias']
    m, v = W[name + '.running_mean'], W[name + '.running_var']
    scale = g / torch.sqrt(v + eps)
    return x * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)


def _norm(x, W, name, kind):
    return instance_norm(x) if kind == 'instance' else batch_norm_eval(x, W, name)


def residual_block(x, W, p, kind, stride):
    """extractor.py:48-56."""
    y = torch.relu(_norm(_conv(x, W, p + '.conv1', stride=stride, padding=1), W, p + '.norm1', kind))
    y = torch.relu(_norm(_conv(y, W, p + '.conv2', padding=1), W, p + '.norm2', kind))
    if stride != 1:
        x = _norm(_conv(x, W, p + '.downsample.0', stride=stride), W, p + '.norm3', kind)
    return torch.relu(x + y)


def basic_encoder(x, W, net):
    """BasicEncoder.forward (extractor.py:168-195).  net='fnet' -> instance norm,
    net='cnet' -> batch norm in eval mode.  x: (B,3,H,W) in [-1,1]."""
    kind = 'instance' if net == 'fnet' else 'batch'
    x = torch.relu(_norm(_conv(x, W, f'{net}.conv1', stride=2, padding=3), W, f'{net}.norm1', kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = residual_block(x, W, f'{net}.layer{li}.0', kind, stride)
        x = residual_block(x, W, f'{net}.layer{li}.1', kind, 1)
    return _conv(x, W, f'{net}.conv2')


# ----------------------------------------------------------------------------------------------
# correlation volume + lookup (MFT/RAFT/core/corr.py:14-69, core/utils/utils.py:98-112)
# ----------------------------------------------------------------------------------------------

def corr_pyramid(fmap1, fmap2, levels=4):
    """All-pairs correlation / sqrt(C), then 2x2 average pooling over the TARGET dims
    (corr.py:20-28,53-69).  fmap: (1,C,h,w).  Returns list of (N, h_l, w_l) tensors."""
    _, C, h, w = fmap1.shape
    a = fmap1.reshape(C, h * w)
    b = fmap2.reshape(C, h * w)
    corr = (a.t() @ b) / math.sqrt(C)
    corr = corr.reshape(h * w, 1, h, w)
    pyr = [corr[:, 0]]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr[:, 0])
    return pyr


def _unnormalize_roundtrip(c, size, via_mul):
    """The reference feeds pixel coordinates through a normalise -> grid_sample
    (align_corners=True) round trip.  ``via_mul`` picks which of its two normalisers:
    core/utils/utils.py:102-103 computes 2*x/(W-1)-1; MFT/utils/interpolation.py:69-72
    computes x*(2/(W-1))-1 with the scale rounded to fp32 first.  ATen then undoes it as
    ((g+1)/2)*(W-1).  All in fp32, one rounding per operation."""
    c = c.astype(F32)
    if via_mul:
        g = c * F32(2.0 / (size - 1)) - F32(1)
    else:
        g = (F32(2) * c) / F32(size - 1) - F32(1)
    return ((g + F32(1)) / F32(2)) * F32(size - 1)


def bilinear_zero(img, px, py, via_mul):
    """Bilinear sample with zero padding, align_corners=True (== F.grid_sample defaults
    used at corr.py:47 via utils.py:106 and results.py:109,133).

    img: (C,H,W) float32 numpy; px,py: arrays of pixel coordinates (any shape S).
    Returns (C,)+S.  Defined operation order (mirrored by the CUDA kernels):
        wE = ix - floor(ix); wW = 1 - wE; wS = iy - floor(iy); wN = 1 - wS
        out = ((wW*wN)*v_nw + (wE*wN)*v_ne) + (wW*wS)*v_sw) + (wE*wS)*v_se
    """
    C, H, W = img.shape
    ix = _unnormalize_roundtrip(px, W, via_mul)
    iy = _unnormalize_roundtrip(py, H, via_mul)
    x0f = np.floor(ix)
    y0f = np.floor(iy)
    wE = (ix - x0f).astype(F32)
    wW = (F32(1) - wE).astype(F32)
    wS = (iy - y0f).astype(F32)
    wN = (F32(1) - wS).astype(F32)
    # clip before the int cast so wild coordinates cannot overflow
    x0 = np.clip(x0f, -2, W + 1).astype(np.int64)
    y0 = np.clip(y0f, -2, H + 1).astype(np.int64)

    def tap(xi, yi):
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        v = img[:, np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)]
        return np.where(ok[None], v, F32(0)).astype(F32)

    nonfinite = ~(np.isfinite(ix) & np.isfinite(iy))
    with np.errstate(invalid='ignore', over='ignore'):
        out = (wW * wN) * tap(x0, y0)
        out = out + (wE * wN) * tap(x0 + 1, y0)
        out = out + (wW * wS) * tap(x0, y0 + 1)
        out = out + (wE * wS) * tap(x0 + 1, y0 + 1)
    out = out.astype(F32)
    if nonfinite.any():
        out = np.where(nonfinite[None], F32(np.nan), out)
    return out


def corr_lookup(pyr, coords, radius=4):
    """CorrBlock.__call__ (corr.py:30-51).  coords: (2,h,w) torch (x,y) at level 0.
    Output (324,h,w): channel = level*81 + (dx+r)*(2r+1) + (dy+r)  -- the reference's
    meshgrid(dy,dx) stacked onto (x,y) makes the FIRST window index the x offset."""
    _, h, w = coords.shape
    n = h * w
    cx = coords[0].reshape(n).numpy().astype(F32)
    cy = coords[1].reshape(n).numpy().astype(F32)
    d = np.arange(-radius, radius + 1, dtype=F32)
    out = []
    rows = np.arange(n)
    for lvl, c in enumerate(pyr):
        c = c.numpy()
        _, hl, wl = c.shape
        sx = (cx / F32(2 ** lvl)).astype(F32)[:, None, None] + d[None, :, None]   # (n,9,1)  x varies with FIRST index
        sy = (cy / F32(2 ** lvl)).astype(F32)[:, None, None] + d[None, None, :]   # (n,1,9)
        sx = np.broadcast_to(sx, (n, 9, 9))
        sy = np.broadcast_to(sy, (n, 9, 9))
        ix = _unnormalize_roundtrip(sx, wl, via_mul=False)
        iy = _unnormalize_roundtrip(sy, hl, via_mul=False)
        x0f, y0f = np.floor(ix), np.floor(iy)
        wE = (ix - x0f).astype(F32); wW = (F32(1) - wE).astype(F32)
        wS = (iy - y0f).astype(F32); wN = (F32(1) - wS).astype(F32)
        x0 = np.clip(x0f, -2, wl + 1).astype(np.int64)
        y0 = np.clip(y0f, -2, hl + 1).astype(np.int64)

        def tap(xi, yi):
            ok = (xi >= 0) & (xi < wl) & (yi >= 0) & (yi < hl)
            v = c[rows[:, None, None], np.clip(yi, 0, hl - 1), np.clip(xi, 0, wl - 1)]
            return np.where(ok, v, F32(0)).astype(F32)

        o = (wW * wN) * tap(x0, y0)
        o = o + (wE * wN) * tap(x0 + 1, y0)
        o = o + (wW * wS) * tap(x0, y0 + 1)
        o = o + (wE * wS) * tap(x0 + 1, y0 + 1)
        out.append(o.astype(F32).reshape(n, 81))
    out = np.concatenate(out, axis=1)                    # (n, 324)
    return torch.from_numpy(np.ascontiguousarray(out.T.reshape(324, h, w)))


# ----------------------------------------------------------------------------------------------
# update block, heads, upsampling (core/update.py, core/raft.py:83-94)
# ----------------------------------------------------------------------------------------------

def motion_encoder(W, flow, corr):
    """BasicMotionEncoder.forward (update.py:152-160)."""
    p = 'update_block.encoder'
    cor = torch.relu(_conv(corr, W, p + '.convc1'))
    cor = torch.relu(_conv(cor, W, p + '.convc2', padding=1))
    flo = torch.relu(_conv(flow, W, p + '.convf1', padding=3))
    flo = torch.relu(_conv(flo, W, p + '.convf2', padding=1))
    out = torch.relu(_conv(torch.cat([cor, flo], 1), W, p + '.conv', padding=1))
    return torch.cat([out, flow], 1)


def sep_conv_gru(W, h, x):
    """SepConvGRU.forward (update.py:108-123): horizontal (1x5) then vertical (5x1)."""
    p = 'update_block.gru'
    for sfx, pad in (('1', (0, 2)), ('2', (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(_conv(hx, W, f'{p}.convz{sfx}', padding=pad))
        r = torch.sigmoid(_conv(hx, W, f'{p}.convr{sfx}', padding=pad))
        q = torch.tanh(_conv(torch.cat([r * h, x], 1), W, f'{p}.convq{sfx}', padding=pad))
        h = (1 - z) * h + z * q
    return h


def flow_head(W, net):
    p = 'update_block.flow_head'
    return _conv(torch.relu(_conv(net, W, p + '.conv1', padding=1)), W, p + '.conv2', padding=1)


def mask_head(W, net):
    """update.py:224-227,237 (0.25 scale)."""
    p = 'update_block.mask'
    return 0.25 * _conv(torch.relu(_conv(net, W, p + '.0', padding=1)), W, p + '.2')


def ou_block(W, net, inp, corr, flow, delta_flow, motion):
    """OcclusionAndUncertaintyBlock.forward (update.py:196-214); concat order :197."""
    x = torch.cat([net, inp, corr, flow, delta_flow, motion], 1)
    outs = []
    for head in ('occl_head', 'uncertainty_head'):
        p = f'occlusion_block.{head}'
        outs.append(_conv(torch.relu(_conv(x, W, p + '.conv1', padding=1)), W, p + '.conv2', padding=1))
    return outs[0], outs[1]


def convex_upsample(x, mask, mult):
    """RAFT.upsample_flow (core/raft.py:83-94).  x: (1,C,h,w), mask: (1,576,h,w) laid out
    (9, 8, 8) = (ky*3+kx, sub_y, sub_x).  out[c, 8y+sy, 8x+sx] =
    sum_k softmax_k(mask[k,sy,sx,y,x]) * (mult*x)[c, y+ky-1, x+kx-1] (zero outside)."""
    _, C, h, w = x.shape
    m = torch.softmax(mask.reshape(9, 8, 8, h, w), dim=0)
    xp = F.pad(mult * x[0], (1, 1, 1, 1))
    out = torch.zeros(C, 8, 8, h, w)
    for k in range(9):
        ky, kx = divmod(k, 3)
        out += m[k][None] * xp[:, ky:ky + h, kx:kx + w][:, None, None]
    return out.permute(0, 3, 1, 4, 2).reshape(1, C, 8 * h, 8 * w)


def coords_grid(h, w):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    return torch.stack([xs, ys], 0).float()


def encode_frame(W, image, want_context=True):
    """Per-frame, pair-independent part of RAFT.forward (core/raft.py:122-149): normalise,
    fnet, cnet -> (fmap, net0, inp).  image: (1,3,H,W) float RGB in [0,255].  The reference
    runs cnet on image1 only (want_context=False for image2)."""
    x = 2 * (image / 255.0) - 1.0
    fmap = basic_encoder(x, W, 'fnet')
    if not want_context:
        return fmap, None, None
    c = basic_encoder(x, W, 'cnet')
    return fmap, torch.tanh(c[:, :128]), torch.relu(c[:, 128:])


def corr_lookup_grid_sample(pyr, coords, radius=4):
    """Same lookup through F.grid_sample, the way the reference itself does it (corr.py:30-51,
    utils.py:98-106).  Used for CPU-baseline TIMING (ATen's vectorised sampler is what the
    reference pays for); tests check it against the explicit corr_lookup above."""
    _, h, w = coords.shape
    n = h * w
    c = coords.permute(1, 2, 0).reshape(n, 1, 1, 2)
    d = torch.linspace(-radius, radius, 2 * radius + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing='ij'), dim=-1).view(1, 2 * radius + 1, 2 * radius + 1, 2)
    out = []
    for lvl, corr in enumerate(pyr):
        hl, wl = corr.shape[-2:]
        pos = c / 2 ** lvl + delta
        gx = 2 * pos[..., 0] / (wl - 1) - 1
        gy = 2 * pos[..., 1] / (hl - 1) - 1
        s = F.grid_sample(corr[:, None], torch.stack([gx, gy], -1), align_corners=True)
        out.append(s.view(n, -1))
    return torch.cat(out, 1).t().reshape(-1, h, w).contiguous()


def raft_forward(W, image1, image2, iters=12, taps=None, fast_lookup=False, flow_init=None):
    """RAFT.forward(test_mode=True) (core/raft.py:97-259).  Images (1,3,H,W) float RGB in
    [0,255], H,W multiples of 8.  Returns dict(flow (1,2,H,W), occlusion logits (1,2,H,W),
    uncertainty (1,1,H,W), coords (1,2,h,w)).  ``taps``: optional dict that receives stage
    boundaries for kernel-level tests."""
    fmap1, net, inp = encode_frame(W, image1)
    fmap2, _, _ = encode_frame(W, image2, want_context=False)
    pyr = corr_pyramid(fmap1, fmap2)
    lookup = corr_lookup_grid_sample if fast_lookup else corr_lookup
    _, _, h, w = fmap1.shape
    coords0 = coords_grid(h, w)
    coords1 = coords0.clone()
    if flow_init is not None:                 # core/raft.py:153-154: (1,2,h,w) coarse flow added to the start coordinates
        coords1 = coords1 + flow_init[0]
    if taps is not None:
        taps.update(fmap1=fmap1, fmap2=fmap2, net0=net, inp=inp, pyramid=pyr, iters=[])
    for itr in range(iters):
        corr = lookup(pyr, coords1)[None]
        flow = (coords1 - coords0)[None]
        motion = motion_encoder(W, flow, corr)
        net = sep_conv_gru(W, net, torch.cat([inp, motion], 1))
        delta = flow_head(W, net)
        coords1 = coords1 + delta[0]
        if taps is not None:
            taps['iters'].append(dict(corr=corr, motion=motion, net=net, delta=delta,
                                      coords1=coords1.clone()))
    mask = mask_head(W, net)
    flow_lo = (coords1 - coords0)[None]
    occ, unc = ou_block(W, net, inp, corr, flow_lo, delta, motion)
    out = dict(flow=convex_upsample(flow_lo, mask, 8.0),
               occlusion=convex_upsample(occ, mask, 1.0),
               uncertainty=convex_upsample(unc, mask, 1.0),
               coords=flow_lo)
    if taps is not None:
        taps.update(mask=mask, occ_lo=occ, unc_lo=unc)
    return out


# ----------------------------------------------------------------------------------------------
# flow wrapper (MFT/raft.py:30-73) and padding (core/utils/utils.py:7-24)
# ----------------------------------------------------------------------------------------------

def pad_amounts(H, W):
    """InputPadder 'sintel' mode: (left, right, top, bottom)."""
    ph = (8 - H % 8) % 8
    pw = (8 - W % 8) % 8
    return pw // 2, pw - pw // 2, ph // 2, ph - ph // 2


def bgr_to_input(img_bgr):
    """uint8 BGR HWC -> float RGB (1,3,H,W), replicate-padded to /8 (raft.py:41-48)."""
    x = torch.from_numpy(np.ascontiguousarray(img_bgr[:, :, ::-1])).permute(2, 0, 1)[None].float()
    l, r, t, b = pad_amounts(*img_bgr.shape[:2])
    if l or r or t or b:
        x = F.pad(x, (l, r, t, b), mode='replicate')
    return x


def postprocess(out, H, W):
    """Unpad, occlusion = softmax(logits)[1], sigma = sqrt(exp(u)) (raft.py:56-62)."""
    l, r, t, b = pad_amounts(H, W)
    sl = (slice(None), slice(None), slice(t, t + H), slice(l, l + W))
    flow = out['flow'][sl][0]
    occ = torch.softmax(out['occlusion'], dim=1)[:, 1:2][sl][0]
    sigma = torch.sqrt(torch.exp(out['uncertainty'][sl][0]))
    return flow, occ, sigma


def downsample_flow_8(flow):
    """MFT/raft.py:98-101: bilinear (align_corners=True) to 1/8 resolution, values / 8."""
    return F.interpolate(flow, size=(flow.shape[2] // 8, flow.shape[3] // 8), mode='bilinear', align_corners=True) / 8


def compute_flow(W, src_bgr, dst_bgr, iters=12, taps=None, fast_lookup=False, init_flow=None):
    """RAFTWrapper.compute_flow(mode='flow') -> flow (2,H,W), occlusion (1,H,W), sigma (1,H,W).
    init_flow: optional (2,H,W) flow, replicate-padded like the images and downsampled by 8 (MFT/raft.py:49-53)."""
    H, Wd = src_bgr.shape[:2]
    flow_init = None
    if init_flow is not None:
        f = torch.as_tensor(init_flow, dtype=torch.float32)[None]
        l, r, t, b = pad_amounts(H, Wd)
        if l or r or t or b:
            f = F.pad(f, (l, r, t, b), mode='replicate')
        flow_init = downsample_flow_8(f)
    out = raft_forward(W, bgr_to_input(src_bgr), bgr_to_input(dst_bgr), iters=iters, taps=taps,
                       fast_lookup=fast_lookup, flow_init=flow_init)
    return postprocess(out, H, Wd)


# ----------------------------------------------------------------------------------------------
# chaining + selection (MFT/MFT.py:114-142,233-239; MFT/results.py:87-136,250-265)
# ----------------------------------------------------------------------------------------------

def chain(left, right):
    """chain_results (MFT.py:233-239).  left/right: tuples (flow (2,H,W), occ (1,H,W),
    sigma (1,H,W)) float32 numpy.  Defined fp32 operation order:
        p      = grid + left.flow
        S      = bilinear_zero(right, p)              (via_mul normaliser)
        flow   = (p + S.flow) - grid                  (results.py:112)
        occ    = max(left.occ, S.occ)
        sigma  = sqrt(left.sigma*left.sigma + S.sigma*S.sigma)
    """
    lf, lo, ls = (np.asarray(a, dtype=F32) for a in left)
    rf, ro, rs = (np.asarray(a, dtype=F32) for a in right)
    _, H, W = lf.shape
    gx = np.broadcast_to(np.arange(W, dtype=F32)[None, :], (H, W))
    gy = np.broadcast_to(np.arange(H, dtype=F32)[:, None], (H, W))
    px = (gx + lf[0]).astype(F32)
    py = (gy + lf[1]).astype(F32)
    S = bilinear_zero(np.concatenate([rf, ro, rs], 0), px, py, via_mul=True)
    fx = ((px + S[0]).astype(F32) - gx).astype(F32)
    fy = ((py + S[1]).astype(F32) - gy).astype(F32)
    occ = np.maximum(lo[0], S[2]).astype(F32)
    # torch.maximum propagates NaN; np.maximum does too.
    with np.errstate(over="ignore", invalid="ignore"):
        sig = np.sqrt((ls[0] * ls[0]).astype(F32) + (S[3] * S[3]).astype(F32)).astype(F32)
    return np.stack([fx, fy]), occ[None], sig[None]


def select(cands, occlusion_threshold):
    """Selection block (MFT.py:114-142) + invalid_mask (results.py:250-265).

    cands: list of chained (flow, occ, sigma) ordered [inf, then ascending delta].
    score = -sigma, -inf where occ > thr; argmax over candidates with torch.max semantics:
    first maximum wins, a NaN score wins over everything (first NaN).  Then occlusion := 1
    where grid + flow leaves [0,W) x [0,H).  Returns (flow, occ, sigma, index uint8 (H,W))."""
    flows = np.stack([c[0] for c in cands]).astype(F32)        # (K,2,H,W)
    occs = np.stack([c[1][0] for c in cands]).astype(F32)      # (K,H,W)
    sigs = np.stack([c[2][0] for c in cands]).astype(F32)
    K, H, W = occs.shape
    score = -sigs
    score[occs > F32(occlusion_threshold)] = -np.inf
    best = np.zeros((H, W), dtype=np.int64)
    bval = score[0].copy()
    for k in range(1, K):
        take = (score[k] > bval) | (np.isnan(score[k]) & ~np.isnan(bval))
        best[take] = k
        bval = np.where(take, score[k], bval)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    flow = flows[best, :, ii, jj].transpose(2, 0, 1).copy()
    occ = occs[best, ii, jj].copy()
    sig = sigs[best, ii, jj].copy()
    gx = np.arange(W, dtype=F32)[None, :]
    gy = np.arange(H, dtype=F32)[:, None]
    ex = (gx + flow[0]).astype(F32)
    ey = (gy + flow[1]).astype(F32)
    invalid = (ex < 0) | (ey < 0) | (ex >= W) | (ey >= H)
    occ[invalid] = F32(1)
    return flow, occ[None], sig[None], best.astype(np.uint8)


def live_chains(deltas, current, start, direction):
    """Delta loop bookkeeping (MFT.py:74-91): which (delta, left_id) pairs are live for
    frame ``current``; then the selection order of MFT.py:114 (inf first, ascending)."""
    used, live = [], []
    for d in deltas:
        if np.isinf(d):
            left = start
        else:
            left = current - int(d) * direction
            before = left < start if direction > 0 else left > start
            if before:
                continue
        if left in used:
            continue
        used.append(left)
        live.append((d, int(left)))
    live.sort(key=lambda t: 0 if np.isinf(t[0]) else t[0])
    return live


class OracleTracker:
    """CPU mirror of MFT.MFT (MFT/MFT.py:13-185) on top of the functions above."""

    def __init__(self, W, deltas=(np.inf, 1, 2, 4, 8, 16, 32), occlusion_threshold=0.02, iters=12,
                 fast_lookup=False):
        self.W, self.deltas, self.thr, self.iters = W, list(deltas), occlusion_threshold, iters
        self.fast_lookup = fast_lookup

    def init(self, img, start_frame_i=0, time_direction=1):
        assert time_direction in (1, -1)
        H, Wd = img.shape[:2]
        self.start, self.cur, self.dir = start_frame_i, start_frame_i, time_direction
        z = (np.zeros((2, H, Wd), F32), np.zeros((1, H, Wd), F32), np.zeros((1, H, Wd), F32))
        self.memory = {start_frame_i: dict(img=img, result=z)}
        return SimpleNamespace(result=z)

    def track(self, img):
        self.cur += self.dir
        live = live_chains(self.deltas, self.cur, self.start, self.dir)
        cands = []
        for _, left in live:
            with torch.no_grad():
                f, o, s = compute_flow(self.W, self.memory[left]['img'], img, self.iters,
                                       fast_lookup=self.fast_lookup)
            cands.append(chain(self.memory[left]['result'], (f.numpy(), o.numpy(), s.numpy())))
        flow, occ, sig, idx = select(cands, self.thr)
        self.memory[self.cur] = dict(img=img, result=(flow, occ, sig))
        finite = [d for d in self.deltas if np.isfinite(d)]
        maxd = max(finite) if finite else 0
        keep_start = any(np.isinf(d) for d in self.deltas)
        for k in list(self.memory):                         # cleanup_memory (MFT.py:157-181)
            if k == self.start and keep_start:
                continue
            if (self.dir > 0 and k + maxd > self.cur) or (self.dir < 0 and k - maxd < self.cur):
                continue
            del self.memory[k]
        return SimpleNamespace(result=(flow, occ, sig), index=idx, live=live)


# ------------------------------------------------------------------------------------------------------------------
# forward splat: FlowOUTrackingResult.warp_forward -> interpolation.bilinear_splat
# ------------------------------------------------------------------------------------------------------------------
def bilinear_splat(data, coords, H, W):
    """MFT/utils/interpolation.py:234-309.  data (N,C) float32, coords (N,2) xy float32 -> accum (H,W,C), counts (H,W).
    Points are NOT dropped at the border: coordinates and the four corner indices are clamped into the grid (x1 is taken
    from the unclamped floor, :256-268), which gives out-of-grid points zero weight on the clamped axis."""
    data = np.asarray(data, np.float32)
    x = np.asarray(coords[:, 0], np.float32)
    y = np.asarray(coords[:, 1], np.float32)
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    x = np.clip(x, 0, W - 1).astype(np.float32)
    y = np.clip(y, 0, H - 1).astype(np.float32)
    x0, x1 = np.clip(x0, 0, W - 1), np.clip(x1, 0, W - 1)
    y0, y1 = np.clip(y0, 0, H - 1), np.clip(y1, 0, H - 1)
    x0f, x1f, y0f, y1f = (v.astype(np.float32) for v in (x0, x1, y0, y1))
    w_a = (x1f - x) * (y1f - y)
    w_b = (x1f - x) * (y - y0f)
    w_c = (x - x0f) * (y1f - y)
    w_d = (x - x0f) * (y - y0f)
    accum = np.zeros((H * W, data.shape[1]), np.float32)
    counts = np.zeros((H * W,), np.float32)
    for wgt, yy, xx in ((w_a, y0, x0), (w_b, y1, x0), (w_c, y0, x1), (w_d, y1, x1)):     # (:289-292 ordering)
        lin = yy * W + xx
        np.add.at(accum, lin, data * wgt[:, None])
        np.add.at(counts, lin, wgt)
    return accum.reshape(H, W, -1), counts.reshape(H, W)


def warp_forward(flow, img, mask=None, border=None):
    """MFT/results.py:190-248.  flow (2,H,W), img (H,W,C) -> (H,W,C): values splatted to grid + flow, normalised by the
    accumulated weight where it is > 0; elsewhere 0 (or `border`)."""
    flow = np.asarray(flow, np.float32)
    H, W = flow.shape[1:]
    img = np.asarray(img, np.float32).reshape(H, W, -1)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    pos = np.stack([xs + flow[0], ys + flow[1]], -1).reshape(-1, 2).astype(np.float32)
    vals = img.reshape(H * W, -1)
    if mask is not None:
        m = np.asarray(mask).reshape(-1).astype(bool)
        pos, vals = pos[m], vals[m]
    accum, counts = bilinear_splat(vals, pos, H, W)
    out = accum.copy()
    nz = counts > 0
    out[nz] /= counts[nz][:, None]
    if border is not None:
        out[~nz] = border
    return out
