"""Python handle on the native engine (libmft_b200.so).  torch is used for device memory and
streams only; every computation is a kernel of the library."""
import ctypes as C

import numpy as np
import torch

from . import _lib, weights as _weights


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One engine = one set of uploaded weights + one frame geometry on the current device."""

    def __init__(self, state_dict):
        if not torch.cuda.is_available():
            raise _lib.MftB200Error('mft_b200 needs a CUDA device (B200); there is no CPU fallback')
        self.L = _lib.lib()
        ctx = C.c_void_p()
        _lib.check(self.L.mftb200_create(C.byref(ctx)))
        self.ctx = ctx
        self.geometry = None
        W = _weights.strip_module_prefix(state_dict)
        for i, (w16, bias, cout_pad, ktot, bias_len) in enumerate(_weights.pack_all(W)):
            w16 = np.ascontiguousarray(w16)
            bias = np.ascontiguousarray(bias)
            _lib.check(self.L.mftb200_upload_layer(self.ctx, i, w16.ctypes.data, bias.ctypes.data, cout_pad, ktot,
                                                   bias_len), self.ctx)

    def __del__(self):
        try:
            if getattr(self, 'ctx', None):
                self.L.mftb200_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    def configure(self, H, W, max_pairs=7, n_slots=34, iters=12):
        geo = (H, W, max_pairs, n_slots, iters)
        if self.geometry != geo:
            _lib.check(self.L.mftb200_configure(self.ctx, H, W, max_pairs, n_slots, iters), self.ctx)
            self.geometry = geo
            self._pinned = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        return self

    def set_option(self, key, value):
        _lib.check(self.L.mftb200_set_option(self.ctx, key.encode(), int(value)), self.ctx)

    def encode_frame(self, frame, slot):
        """frame: (H,W,3) uint8 BGR numpy array (host) or torch CUDA tensor.

        A numpy frame in page-locked memory is DMA'd straight from the caller's buffer: the caller must leave it
        unchanged until ``wait_frame_copied()`` returns (or until a result that depends on the frame is back on the
        host).  Returns True when that in-place path was taken."""
        H, W = self.geometry[:2]
        if isinstance(frame, torch.Tensor) and frame.is_cuda:
            assert frame.dtype == torch.uint8 and tuple(frame.shape) == (H, W, 3) and frame.is_contiguous()
            _lib.check(self.L.mftb200_encode_frame(self.ctx, C.c_void_p(frame.data_ptr()), 1, slot, _stream_ptr()), self.ctx)
            return False
        frame = np.asarray(frame)
        assert frame.dtype == np.uint8 and frame.shape == (H, W, 3), (frame.dtype, frame.shape)
        if frame.flags.c_contiguous and self.L.mftb200_is_pinned_host(C.c_void_p(frame.ctypes.data)):
            # the caller's frame is page-locked: DMA straight from it (the caller keeps it unchanged until the result of this
            # frame is back, like the reference, whose memory holds references to the caller's frames, MFT.py:150)
            _lib.check(self.L.mftb200_encode_frame(self.ctx, C.c_void_p(frame.ctypes.data), 0, slot, _stream_ptr()), self.ctx)
            return True
        # stage through pinned memory so the H2D copy is asynchronous w.r.t. the host
        torch.cuda.current_stream().synchronize()
        self._pinned.numpy()[...] = frame
        _lib.check(self.L.mftb200_encode_frame(self.ctx, C.c_void_p(self._pinned.data_ptr()), 0, slot, _stream_ptr()), self.ctx)
        return False

    def refine(self, left_slots, right_slots, out=None, init_flow=None):
        """Batched RAFT-OU refinement.  Returns CUDA float tensor (n_pairs, 4, H, W).
        init_flow: optional CUDA float (n_pairs, 2, h, w) coarse flow added to the start coordinates (RAFT's flow_init)."""
        H, W = self.geometry[:2]
        n = len(left_slots)
        assert n == len(right_slots) and 1 <= n <= self.geometry[2]
        if out is None:
            out = torch.empty((n, 4, H, W), dtype=torch.float32, device='cuda')
        ls = (C.c_int * n)(*[int(s) for s in left_slots])
        rs = (C.c_int * n)(*[int(s) for s in right_slots])
        if init_flow is None:
            _lib.check(self.L.mftb200_raft_refine(self.ctx, n, ls, rs, C.c_void_p(out.data_ptr()), _stream_ptr()), self.ctx)
        else:
            hp, wp = (H + 7) // 8, (W + 7) // 8
            assert init_flow.is_cuda and init_flow.dtype == torch.float32 and init_flow.is_contiguous() and tuple(init_flow.shape) == (n, 2, hp, wp)
            _lib.check(self.L.mftb200_raft_refine_init(self.ctx, n, ls, rs, C.c_void_p(init_flow.data_ptr()), C.c_void_p(out.data_ptr()),
                                                       _stream_ptr()), self.ctx)
        return out

    def slot_tensors(self):
        """Zero-copy torch views of the feature-slot arrays (fmap fp16 [S,N,256], net fp32 [S,N,128], inp fp16 [S,N,128]) for
        multi-GPU feature exchange; work enqueued on the current stream afterwards sees all earlier encodes complete."""
        ptrs = [C.c_void_p() for _ in range(3)]
        nbytes = (C.c_size_t * 3)()
        _lib.check(self.L.mftb200_slot_buffers(self.ctx, C.byref(ptrs[0]), C.byref(ptrs[1]), C.byref(ptrs[2]), nbytes, _stream_ptr()), self.ctx)
        S = self.geometry[3]
        out = []
        for p, nb, (ch, dt, ts) in zip(ptrs, nbytes, ((256, torch.float16, '<f2'), (128, torch.float32, '<f4'), (128, torch.float16, '<f2'))):
            n = nb // (ch * (2 if dt == torch.float16 else 4))

            class _Mem:          # __cuda_array_interface__ carrier: torch.as_tensor wraps the memory without a copy
                pass
            m = _Mem()
            m.__cuda_array_interface__ = {'shape': (S, n, ch), 'typestr': ts, 'data': (p.value, False), 'version': 2, 'strides': None}
            t = torch.as_tensor(m, device='cuda')
            t._mftb200_owner = self       # the engine owns the memory
            out.append(t)
        return out

    def check_device(self):
        """Synchronises the device and raises if a kernel reported a pipeline time-out."""
        self._checked(self.L.mftb200_device_error_flag(self.ctx))

    def error_flag_async(self):
        """Enqueues a copy of the device error flag into its pinned mirror on the current stream (no sync)."""
        _lib.check(self.L.mftb200_error_flag_async(self.ctx, _stream_ptr()), self.ctx)

    def error_flag_poll(self):
        """Raises if the mirror shows an aborted kernel (as of the last error_flag_async the stream has completed)."""
        self._checked(self.L.mftb200_error_flag_poll(self.ctx))

    def wait_frame_copied(self):
        """Blocks until the newest encode_frame's host->device copy has left the caller's buffer."""
        _lib.check(self.L.mftb200_wait_frame_copied(self.ctx), self.ctx)

    def _checked(self, code):
        if code != 0:
            # an aborted launch leaves the persistent kernels' work queues undefined: the library refuses further launches
            # until the workspace is rebuilt, so the next configure() must not be skipped as "same geometry"
            self.geometry = None
        _lib.check(code, self.ctx)

    def profile_fetch(self):
        """(ms, steps) per kind: index 0 = tensor-core conv launches, 1 = bandwidth-bound kernels."""
        ms = (C.c_double * 2)()
        n = (C.c_longlong * 2)()
        _lib.check(self.L.mftb200_profile_fetch(self.ctx, ms, n), self.ctx)
        return [float(ms[0]), float(ms[1])], [int(n[0]), int(n[1])]

    def profile_steps(self, max_steps=4096):
        """Per-step (ms, kind, layer) in launch order since profiling was switched on; kind 0 = tensor-core
        conv/GEMM, 1 = bandwidth-bound kernel(s); layer = index into weights.LAYER_NAMES, 100 = correlation
        GEMM, -1 = not a conv."""
        ms = (C.c_float * max_steps)()
        kinds = (C.c_int * max_steps)()
        n = C.c_int(0)
        _lib.check(self.L.mftb200_profile_steps(self.ctx, ms, kinds, max_steps, C.byref(n)), self.ctx)
        return [(float(ms[i]), int(kinds[i]) & 0xff, (int(kinds[i]) >> 8) - 1) for i in range(min(n.value, max_steps))]

    def launch_count(self):
        return int(self.L.mftb200_launch_count(self.ctx))

    def debug_buffer(self, name, dtype, shape):
        """Copy of (the head of) a named internal buffer as a torch CUDA tensor."""
        n = int(np.prod(shape))
        out = torch.empty(n, dtype=dtype, device='cuda')
        torch.cuda.synchronize()
        _lib.check(self.L.mftb200_debug_read(self.ctx, name.encode(), C.c_void_p(out.data_ptr()),
                                             n * out.element_size()), self.ctx)
        return out.reshape(shape)


def chain_select(lefts, right, occlusion_threshold, want_index=True):
    """Fused flow-chain composition + per-pixel best-chain selection + invalid mask.

    lefts: list of K CUDA float tensors (4,H,W) (template -> left_k), right: (K,4,H,W)
    (left_k -> current).  Candidate order must be [inf, ascending delta].  Returns
    (result (4,H,W), index uint8 (H,W) or None)."""
    L = _lib.lib()
    K = len(lefts)
    assert right.is_cuda and right.dtype == torch.float32 and right.is_contiguous() and right.shape[0] == K
    _, _, H, W = right.shape
    for t in lefts:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (4, H, W)
    out = torch.empty((4, H, W), dtype=torch.float32, device=right.device)
    idx = torch.empty((H, W), dtype=torch.uint8, device=right.device) if want_index else None
    ptrs = (C.c_void_p * K)(*[t.data_ptr() for t in lefts])
    _lib.check(L.mftb200_chain_select(K, ptrs, C.c_void_p(right.data_ptr()), float(occlusion_threshold), H, W,
                                      C.c_void_p(out.data_ptr()), C.c_void_p(idx.data_ptr() if want_index else None),
                                      _stream_ptr()))
    return out, idx


def set_global_option(key, value):
    _lib.check(_lib.lib().mftb200_set_global_option(key.encode(), int(value)))


def conv2d_bench(x16, w16, bias, cin, cout_pad, n_tile, kh, kw, stride, relu, cluster, smem_cap_kib, reps,
                 timing=None):
    """Tuning aid: the product conv kernel `reps` times; returns (out, mean ms per launch).
    timing: optional CUDA int64 tensor [n_ctas, 8] receiving per-CTA phase timestamps."""
    L = _lib.lib()
    B, H, W, pitch = x16.shape
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    out = torch.zeros((B, Ho, Wo, cout_pad), dtype=torch.float32, device='cuda')
    ms = C.c_float(0)
    _lib.check(L.mftb200_conv2d_bench2(C.c_void_p(x16.data_ptr()), B, H, W, pitch, cin, C.c_void_p(w16.data_ptr()),
                                       C.c_void_p(bias.data_ptr()), cout_pad, n_tile, kh, kw, stride, int(relu),
                                       C.c_void_p(out.data_ptr()), 0, int(cluster), int(smem_cap_kib), int(reps),
                                       C.byref(ms), C.c_void_p(timing.data_ptr() if timing is not None else None),
                                       _stream_ptr()))
    return out, float(ms.value)


def conv2d_test(x16, w16, bias, cin, cout_pad, n_tile, kh, kw, stride, relu, impl):
    """Unit-test hook: x16 CUDA fp16 (B,H,W,pitch); w16 CUDA fp16 [cout_pad][ktot]; returns fp32 (B,Ho,Wo,cout_pad)."""
    L = _lib.lib()
    B, H, W, pitch = x16.shape
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    out = torch.zeros((B, Ho, Wo, cout_pad), dtype=torch.float32, device='cuda')
    _lib.check(L.mftb200_conv2d_test(C.c_void_p(x16.data_ptr()), B, H, W, pitch, cin, C.c_void_p(w16.data_ptr()),
                                     C.c_void_p(bias.data_ptr()), cout_pad, n_tile, kh, kw, stride, int(relu),
                                     C.c_void_p(out.data_ptr()), int(impl), _stream_ptr()))
    return out
