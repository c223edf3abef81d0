"""Checkpoint -> packed tensor-core operands.

Consumes the reference checkpoint's state dict (keys 'module.<name>', loaded at
MFT/raft.py:20-21 of the reference) and produces, per conv layer of the engine's program, the
fp16 "B matrix" [cout_pad][taps * cin_pad64] (K contiguous) plus an fp32 bias vector:

* cnet BatchNorm (eval) is folded into the preceding conv (extractor.py:23-26,120-121);
* convz|convr of each GRU half-step are stacked into one N=256 GEMM; convq's input channels are
  permuted to the engine's [inp | motion | r*h] operand order (update.py:110-113);
* the two OU heads share their 712-channel input: conv1s stacked (N=256), conv2s block-diagonal;
* the 7x7 convs (encoder conv1, convf1) are stored for im2col'ed operands, k = (ky*7+kx)*cin + c.

Layer order == enum Layer in csrc/engine.cu.
"""
import numpy as np
import torch

ENC_LAYERS = ['conv1', 'layer1.0.conv1', 'layer1.0.conv2', 'layer1.1.conv1', 'layer1.1.conv2',
              'layer2.0.conv1', 'layer2.0.conv2', 'layer2.0.downsample.0', 'layer2.1.conv1', 'layer2.1.conv2',
              'layer3.0.conv1', 'layer3.0.conv2', 'layer3.0.downsample.0', 'layer3.1.conv1', 'layer3.1.conv2',
              'conv2']
LAYER_NAMES = ([f'fnet.{n}' for n in ENC_LAYERS] + [f'cnet.{n}' for n in ENC_LAYERS] +
               ['convc1', 'convc2', 'convf1', 'convf2', 'convm', 'gru_zr1', 'gru_q1', 'gru_zr2', 'gru_q2',
                'fh1', 'fh2', 'mask1', 'mask2', 'ou1', 'ou2'])
assert len(LAYER_NAMES) == 47


def strip_module_prefix(state_dict):
    out = {}
    for k, v in state_dict.items():
        k = k[len('module.'):] if k.startswith('module.') else k
        if not k.endswith('num_batches_tracked'):
            out[k] = v.detach().float().cpu()
    return out


def load_checkpoint(path):
    # weights_only: a checkpoint is data from an untrusted source -- never unpickle arbitrary objects
    return strip_module_prefix(torch.load(path, map_location='cpu', weights_only=True))


def _pack(w, b, cout_pad=None):
    """w: (Cout, Cin, kh, kw) fp32 torch, b: (Cout,) -> (uint16 [cout_pad, taps*cin_pad], fp32 bias, ktot)."""
    cout, cin, kh, kw = w.shape
    cin_pad = (cin + 63) // 64 * 64
    cout_pad = cout_pad or (cout + 15) // 16 * 16
    m = torch.zeros(cout_pad, kh * kw, cin_pad, dtype=torch.float32)
    m[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    w16 = m.reshape(cout_pad, kh * kw * cin_pad).to(torch.float16).contiguous().numpy().view(np.uint16)
    bias_len = (cout_pad + 31) // 32 * 32
    bias = np.zeros(bias_len, np.float32)
    bias[:cout] = b.numpy()
    return w16, bias, cout_pad, kh * kw * cin_pad, bias_len


def _im2col_weight(w):
    """(Cout, Cin, kh, kw) -> (Cout, kh*kw*Cin, 1, 1) with k = (ky*kw+kx)*Cin + c."""
    cout = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(cout, -1, 1, 1)


def _fold_bn(W, conv, bn, eps=1e-5):
    w, b = W[conv + '.weight'], W[conv + '.bias']
    g, beta = W[bn + '.weight'], W[bn + '.bias']
    mean, var = W[bn + '.running_mean'], W[bn + '.running_var']
    s = g / torch.sqrt(var + eps)
    return w * s.view(-1, 1, 1, 1), b * s + (beta - mean * s)


def pack_all(W):
    """W: flat state dict (no 'module.' prefix).  Returns list of 47 packed layers."""
    out = []
    for net in ('fnet', 'cnet'):
        for name in ENC_LAYERS:
            conv = f'{net}.{name}'
            if net == 'cnet' and name != 'conv2':
                if name == 'conv1':
                    bn = 'cnet.norm1'
                elif name.endswith('downsample.0'):
                    bn = conv[:-len('downsample.0')] + 'norm3'
                else:
                    bn = conv[:-len('convN')] + 'norm' + name[-1]
                w, b = _fold_bn(W, conv, bn)
            else:
                w, b = W[conv + '.weight'], W[conv + '.bias']
            if name == 'conv1':
                w = _im2col_weight(w)
            out.append(_pack(w, b))
    ub, ob = 'update_block.', 'occlusion_block.'
    g = lambda n: (W[n + '.weight'], W[n + '.bias'])
    out.append(_pack(*g(ub + 'encoder.convc1')))
    out.append(_pack(*g(ub + 'encoder.convc2')))
    wf1, bf1 = g(ub + 'encoder.convf1')
    out.append(_pack(_im2col_weight(wf1), bf1))
    out.append(_pack(*g(ub + 'encoder.convf2')))
    out.append(_pack(*g(ub + 'encoder.conv'), cout_pad=128))
    for sfx in ('1', '2'):
        wz, bz = g(ub + f'gru.convz{sfx}')
        wr, br = g(ub + f'gru.convr{sfx}')
        out.append(_pack(torch.cat([wz, wr], 0), torch.cat([bz, br], 0)))
        wq, bq = g(ub + f'gru.convq{sfx}')
        wq = torch.cat([wq[:, 128:384], wq[:, 0:128]], 1)      # [r*h | inp | motion] -> [inp | motion | r*h]
        out.append(_pack(wq, bq))
    out.append(_pack(*g(ub + 'flow_head.conv1')))
    out.append(_pack(*g(ub + 'flow_head.conv2'), cout_pad=16))
    out.append(_pack(*g(ub + 'mask.0')))
    out.append(_pack(*g(ub + 'mask.2')))
    wo, bo = g(ob + 'occl_head.conv1')
    wu, bu = g(ob + 'uncertainty_head.conv1')
    out.append(_pack(torch.cat([wo, wu], 0), torch.cat([bo, bu], 0)))
    wo2, bo2 = g(ob + 'occl_head.conv2')
    wu2, bu2 = g(ob + 'uncertainty_head.conv2')
    w2 = torch.zeros(3, 256, 3, 3)
    w2[0:2, 0:128] = wo2
    w2[2:3, 128:256] = wu2
    out.append(_pack(w2, torch.cat([bo2, bu2], 0), cout_pad=16))
    assert len(out) == 47
    return out


# ------------------------------------------------------------------------------------------------------------------
# weights without the oracle: checkpoint discovery and a random initialisation of the architecture (benchmarks)
# ------------------------------------------------------------------------------------------------------------------
CKPT_NAME = 'raft-things-sintel-kubric-splitted-occlusion-uncertainty-non-occluded-base-sintel.pth'


def find_checkpoint():
    """Path of the shipped RAFT-OU checkpoint (configs/flow/RAFTou_kubric_huber_split_nonoccl.py:25 of the reference) if it
    is reachable: $MFT_CHECKPOINT, the copy that travels with the repo snapshot, a reference checkout.  None otherwise."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get('MFT_CHECKPOINT', ''), os.path.join(root, 'oracle', '_ref', 'raft_ou_checkpoint.pth'),
             os.path.join(os.environ.get('MFT_REFERENCE_ROOT', '/root/reference'), 'checkpoints', CKPT_NAME),
             os.path.join('checkpoints', CKPT_NAME)]
    for p in cands:
        if p and os.path.isfile(p):
            return p
    return None


def architecture_spec():
    """(name, shape) of every tensor of the shipped RAFT-OU architecture that the path consumes (SURVEY.md Appendix B;
    MFT/RAFT/core/{extractor,update,raft}.py)."""
    spec = []

    def conv(name, co, ci, kh, kw):
        spec.extend([(name + '.weight', (co, ci, kh, kw)), (name + '.bias', (co,))])

    def bn(name, c):
        spec.extend([(f'{name}.{s}', (c,)) for s in ('weight', 'bias', 'running_mean', 'running_var')])

    for net in ('fnet', 'cnet'):
        norm = bn if net == 'cnet' else (lambda name, c: None)
        norm(f'{net}.norm1', 64)
        conv(f'{net}.conv1', 64, 3, 7, 7)
        cin = 64
        for li, dim in ((1, 64), (2, 96), (3, 128)):
            for bi in (0, 1):
                p = f'{net}.layer{li}.{bi}'
                conv(p + '.conv1', dim, cin if bi == 0 else dim, 3, 3)
                conv(p + '.conv2', dim, dim, 3, 3)
                norm(p + '.norm1', dim)
                norm(p + '.norm2', dim)
                if bi == 0 and li > 1:
                    norm(p + '.norm3', dim)
                    conv(p + '.downsample.0', dim, cin, 1, 1)
            cin = dim
        conv(f'{net}.conv2', 256, 128, 1, 1)
    ub, ob = 'update_block.', 'occlusion_block.'
    for name, co, ci, kh, kw in (('encoder.convc1', 256, 324, 1, 1), ('encoder.convc2', 192, 256, 3, 3), ('encoder.convf1', 128, 2, 7, 7),
                                 ('encoder.convf2', 64, 128, 3, 3), ('encoder.conv', 126, 256, 3, 3),
                                 ('gru.convz1', 128, 384, 1, 5), ('gru.convr1', 128, 384, 1, 5), ('gru.convq1', 128, 384, 1, 5),
                                 ('gru.convz2', 128, 384, 5, 1), ('gru.convr2', 128, 384, 5, 1), ('gru.convq2', 128, 384, 5, 1),
                                 ('flow_head.conv1', 256, 128, 3, 3), ('flow_head.conv2', 2, 256, 3, 3), ('mask.0', 256, 128, 3, 3),
                                 ('mask.2', 576, 256, 1, 1)):
        conv(ub + name, co, ci, kh, kw)
    for name, co, ci in (('occl_head.conv1', 128, 712), ('occl_head.conv2', 2, 128), ('uncertainty_head.conv1', 128, 712),
                         ('uncertainty_head.conv2', 1, 128)):
        conv(ob + name, co, ci, 3, 3)
    return spec


def random_init(seed=0):
    """Random weights of the architecture for benchmarks without the checkpoint (fan-in scaled, so activations stay O(1);
    the recurrent / output convolutions a little smaller so the refinement iterations stay bounded)."""
    gen = torch.Generator().manual_seed(int(seed))
    W = {}
    for name, shape in architecture_spec():
        if name.endswith('.weight') and len(shape) == 4:
            gain = (2.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
            gain *= 0.7 if '.gru.' in name else 1.0
            gain *= 0.25 if name.endswith(('flow_head.conv2.weight', 'head.conv2.weight')) else 1.0
            W[name] = torch.randn(shape, generator=gen) * gain
        elif name.endswith('running_var'):
            W[name] = torch.rand(shape, generator=gen) + 0.5
        elif name.endswith('running_mean'):
            W[name] = torch.randn(shape, generator=gen) * 0.1
        elif name.endswith('.weight'):
            W[name] = torch.rand(shape, generator=gen) * 0.4 + 0.8
        else:
            W[name] = torch.randn(shape, generator=gen) * 0.05
    return W
