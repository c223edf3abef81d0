"""FlowOUTrackingResult with the reference's surface (MFT/results.py of serycjon/MFT).

Storage is one planar (4,H,W) fp32 tensor [flow_x, flow_y, occlusion, sigma]; .flow / .occlusion /
.sigma are views of it, so the engine's kernels consume and produce results without copies.
Geometry on CUDA tensors (chain, warp_backward, warp_forward_points, sample) runs in the library's
kernels; the same methods on CPU tensors (what callers hold after track(): meta.result is a CPU
copy, MFT.py:145-148) use torch ops -- that is caller-side convenience, not the tracking path."""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib


def _sptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_f32(t, device=None):
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.asarray(t))
    return t.to(device=device if device is not None else t.device, dtype=torch.float32)


def _normalize(coords, H, W):
    """(..., xy) pixel coordinates -> grid_sample coordinates (MFT/utils/interpolation.py:63-73)."""
    scale = torch.tensor([2 / (W - 1), 2 / (H - 1)], dtype=torch.float32, device=coords.device)
    return coords * scale - 1


class FlowOUTrackingResult:
    def __init__(self, flow, occlusion=None, sigma=None, validate=False):
        """flow (2,H,W), occlusion (1,H,W), sigma (1,H,W).  validate=True re-enables the
        reference's range asserts (results.py:31-33); they force a device->host sync each, so
        the tracker leaves them off."""
        assert flow.dim() == 3 and flow.shape[0] == 2
        self.H, self.W = int(flow.shape[1]), int(flow.shape[2])
        dev = flow.device
        if occlusion is None:
            occlusion = torch.zeros((1, self.H, self.W), dtype=torch.float32, device=dev)
        if sigma is None:
            sigma = torch.zeros((1, self.H, self.W), dtype=torch.float32, device=dev)
        assert tuple(occlusion.shape) == (1, self.H, self.W) and tuple(sigma.shape) == (1, self.H, self.W)
        if validate:
            assert torch.all(occlusion >= 0) and torch.all(occlusion <= 1.000001) and torch.all(sigma >= 0)
        self._packed = None
        self.flow, self.occlusion, self.sigma = flow, occlusion, sigma

    # ---- packed storage ---------------------------------------------------------------------
    @classmethod
    def from_packed(cls, packed):
        assert packed.dim() == 3 and packed.shape[0] == 4 and packed.dtype == torch.float32
        r = cls(packed[0:2], packed[2:3], packed[3:4])
        r._packed = packed
        return r

    def packed(self):
        """Contiguous (4,H,W) [fx, fy, occlusion, sigma] (no copy when built by from_packed)."""
        p = self._packed
        if (p is not None and p.is_contiguous() and self.flow.data_ptr() == p.data_ptr()
                and self.occlusion.data_ptr() == p[2:3].data_ptr() and self.sigma.data_ptr() == p[3:4].data_ptr()):
            return p
        p = torch.cat([self.flow.float(), self.occlusion.float().to(self.flow.device),
                       self.sigma.float().to(self.flow.device)], 0).contiguous()
        return p

    def __repr__(self):
        return f'<{self.__class__.__name__} ({self.H} x {self.W}) has flow, occlusion, sigma>'

    def _rebuild(self, packed):
        self._packed = packed
        self.flow, self.occlusion, self.sigma = packed[0:2], packed[2:3], packed[3:4]
        return self

    def cpu(self):
        return self._rebuild(self.packed().cpu())

    def cuda(self):
        return self._rebuild(self.packed().cuda())

    def clone(self):
        return FlowOUTrackingResult.from_packed(self.packed().clone())

    @classmethod
    def identity(cls, flow_shape, device=None):
        """Zero-flow, zero-sigma, zero-occlusion result (results.py:75-85)."""
        return cls.from_packed(torch.zeros((4, int(flow_shape[0]), int(flow_shape[1])), dtype=torch.float32, device=device))

    # ---- geometry -----------------------------------------------------------------------------
    def _warp(self, img, add_flow):
        img = _as_f32(img, self.flow.device).contiguous()
        assert img.dim() == 3 and tuple(img.shape[1:]) == (self.H, self.W)
        flow = self.flow.float().contiguous()
        if flow.is_cuda:
            out = torch.empty_like(img)
            _lib.check(_lib.lib().mftb200_warp_backward(C.c_void_p(flow.data_ptr()), C.c_void_p(img.data_ptr()),
                                                        int(img.shape[0]), self.H, self.W, int(add_flow),
                                                        C.c_void_p(out.data_ptr()), _sptr()))
            return out
        ys, xs = torch.meshgrid(torch.arange(self.H), torch.arange(self.W), indexing='ij')
        grid = torch.stack([xs, ys], 0).float()
        pos = grid + flow
        samp = F.grid_sample(img[None], _normalize(pos.permute(1, 2, 0)[None], self.H, self.W), align_corners=True)[0]
        return pos + samp - grid if add_flow else samp

    def chain(self, flow):
        """Flow A->C from self (A->B) followed by ``flow`` (B->C) (results.py:87-114)."""
        assert flow.dim() == 3 and flow.shape[0] == 2
        return self._warp(flow, add_flow=True)

    def warp_backward(self, img):
        """Sample img (C,H,W) at the end points of self.flow (results.py:116-136)."""
        return self._warp(img, add_flow=False)

    def invalid_mask(self):
        """(H,W) bool: flow end point outside the image (results.py:250-265)."""
        dev = self.flow.device
        ys, xs = torch.meshgrid(torch.arange(self.H, device=dev), torch.arange(self.W, device=dev), indexing='ij')
        ex, ey = xs + self.flow[0].float(), ys + self.flow[1].float()
        return (ex < 0) | (ey < 0) | (ex >= self.W) | (ey >= self.H)

    def _sample_points(self, field, points, add_points):
        points = _as_f32(points)
        N = int(points.shape[0])
        if field.is_cuda:
            pts = points.to(field.device).contiguous()
            field = field.float().contiguous()
            out = torch.empty((field.shape[0], N), dtype=torch.float32, device=field.device)
            _lib.check(_lib.lib().mftb200_sample_points(C.c_void_p(field.data_ptr()), int(field.shape[0]), self.H, self.W,
                                                        C.c_void_p(pts.data_ptr()), N, int(add_points),
                                                        C.c_void_p(out.data_ptr()), _sptr()))
            return out
        dev = points.device
        g = _normalize(points.view(1, 1, N, 2), self.H, self.W)
        out = F.grid_sample(field.float().to(dev)[None], g, align_corners=True)[0, :, 0]
        if add_points:
            out = out.clone()
            out[:2] += points.t()
        return out

    def warp_forward_points(self, points):
        """(N, xy) source points -> (N, xy) positions in the current frame (results.py:138-157)."""
        return self._sample_points(self.flow, points, add_points=True).t()

    def sample(self, points):
        """Flow (2,N), occlusion (1,N), sigma (1,N) at the query points (results.py:159-188)."""
        s = self._sample_points(self.packed(), points, add_points=False)
        return s[0:2], s[2:3], s[3:4]

    def warp_forward(self, img, mask=None, border=None, numpy_out=True):
        """Forward-splat img (H,W,...) by self.flow with the reference's clamped bilinear weights, normalised by the
        accumulated weight (results.py:190-248 -> interpolation.bilinear_splat, interpolation.py:234-309; demo.py's edit
        propagation).  On a CUDA result this is the library's splat kernel (float atomics); on a CPU result the same
        arithmetic in torch ops (caller-side convenience, not the tracking path)."""
        dev = self.flow.device
        img_t = _as_f32(img, dev)
        H, W = self.H, self.W
        assert tuple(img_t.shape[:2]) == (H, W)
        vals = img_t.reshape(H, W, -1).contiguous()
        Cn = int(vals.shape[2])
        m = None
        if mask is not None:
            m = torch.as_tensor(np.asarray(mask) if not isinstance(mask, torch.Tensor) else mask, device=dev).reshape(H, W)
        if self.flow.is_cuda:
            flow = self.flow.float().contiguous()
            out = torch.empty((H, W, Cn), dtype=torch.float32, device=dev)
            cnt = torch.empty((H, W), dtype=torch.float32, device=dev)
            m8 = m.to(torch.uint8).contiguous() if m is not None else None
            _lib.check(_lib.lib().mftb200_warp_forward(C.c_void_p(flow.data_ptr()), C.c_void_p(vals.data_ptr()),
                                                       C.c_void_p(m8.data_ptr() if m8 is not None else None), Cn, H, W,
                                                       int(border is not None), float(border if border is not None else 0.0),
                                                       C.c_void_p(out.data_ptr()), C.c_void_p(cnt.data_ptr()), _sptr()))
        else:
            ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
            x = (xs + self.flow[0].float()).reshape(-1)
            y = (ys + self.flow[1].float()).reshape(-1)
            v = vals.reshape(H * W, Cn)
            if m is not None:
                keep = m.reshape(-1).bool()
                x, y, v = x[keep], y[keep], v[keep]
            x0, y0 = torch.floor(x).long(), torch.floor(y).long()
            x1, y1 = x0 + 1, y0 + 1
            x, y = x.clamp(0, W - 1), y.clamp(0, H - 1)
            x0, x1, y0, y1 = x0.clamp(0, W - 1), x1.clamp(0, W - 1), y0.clamp(0, H - 1), y1.clamp(0, H - 1)
            acc = torch.zeros((H * W, Cn), dtype=torch.float32)
            cnt = torch.zeros((H * W,), dtype=torch.float32)
            for wgt, yy, xx in (((x1 - x) * (y1 - y), y0, x0), ((x1 - x) * (y - y0), y1, x0),
                                ((x - x0) * (y1 - y), y0, x1), ((x - x0) * (y - y0), y1, x1)):
                lin = yy * W + xx
                acc.index_add_(0, lin, v * wgt[:, None])
                cnt.index_add_(0, lin, wgt)
            nz = cnt > 0
            out = acc.clone()
            out[nz] = out[nz] / cnt[nz][:, None]
            if border is not None:
                out[~nz] = float(border)
            out = out.reshape(H, W, Cn)
        out = out.reshape(img_t.shape)
        return out.cpu().numpy() if numpy_out else out

    # ---- persistence ----------------------------------------------------------------------------
    def write(self, path):
        """numpy .npz with the three fields (the reference's lossy 16-bit codecs are host I/O and
        out of scope)."""
        np.savez_compressed(path, flow=self.flow.cpu().numpy(), occlusion=self.occlusion.cpu().numpy(),
                            sigma=self.sigma.cpu().numpy())

    @classmethod
    def read(cls, path):
        d = np.load(path)
        return cls(torch.from_numpy(d['flow']), torch.from_numpy(d['occlusion']), torch.from_numpy(d['sigma']))
