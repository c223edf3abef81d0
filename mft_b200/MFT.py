"""MFT tracker with the reference's class surface (MFT/MFT.py of serycjon/MFT): MFT(config),
.init(img, start_frame_i, time_direction, flow_cache), .track(img) -> meta with meta.result on CPU.

Per frame (MFT.py:55-154): for every chain delta the flow (t-delta -> t) is chained onto the stored
template->(t-delta) result and, per pixel, the non-occluded candidate with the lowest sigma wins.
Here that is three engine calls: encode the new frame once, ONE batched refinement for all live
delta pairs, ONE fused chain+select kernel."""
import logging
from types import SimpleNamespace

import numpy as np
import torch

from .engine import chain_select
from .raft import MAX_PAIRS, TRACKER_SLOTS
from .results import FlowOUTrackingResult

logger = logging.getLogger(__name__)


class _PinnedPool:
    """Recycled pinned host buffers for the per-frame device->host copy of the result.

    cudaHostAlloc / cudaFreeHost cost milliseconds and synchronise the device, so results are handed out as
    tensors over slots of ONE pinned allocation; a slot returns to the pool when the caller drops the last
    tensor (or view) that references it.  If the caller keeps more results alive than there are slots (demo.py
    keeps every frame) the copy falls back to a plain pageable tensor."""

    def __init__(self, shape, slots=6):
        import weakref
        self._weakref = weakref
        self.base = torch.empty((slots,) + tuple(shape), dtype=torch.float32).pin_memory()
        self.base_np = self.base.numpy()
        self.free = list(range(slots))

    def take(self):
        if not self.free:
            return None
        i = self.free.pop()
        view = self.base_np[i]                       # fresh ndarray object; torch.from_numpy keeps it alive
        self._weakref.finalize(view, self.free.append, i)
        return torch.from_numpy(view)


class MFT:
    def __init__(self, config):
        self.C = config                     # the runner re-assigns tracker.C between runs (run_MFT_tapvid.py:151)
        self.flower = config.flow_config.of_class(config.flow_config)
        self.device = 'cuda'

    # ------------------------------------------------------------------------------------------
    def init(self, img, start_frame_i=0, time_direction=1, flow_cache=None, **kwargs):
        """img: (H,W,3) uint8 BGR.  Returns meta with the zero-motion result on CPU (MFT.py:22-53)."""
        assert time_direction in [+1, -1]
        self.img_H, self.img_W = img.shape[:2]
        self.start_frame_i = start_frame_i
        self.current_frame_i = start_frame_i
        self.time_direction = time_direction
        self.flow_cache = flow_cache
        finite = [d for d in self.C.deltas if np.isfinite(d)]
        if finite and max(finite) + 2 > TRACKER_SLOTS:
            raise ValueError(f'deltas up to {TRACKER_SLOTS - 2} are supported (feature-slot budget)')
        self.engine = self.flower.ensure_geometry(self.img_H, self.img_W, claim=True)
        self._frame_in_place = False
        self._free_slots = list(range(TRACKER_SLOTS))
        if getattr(self, '_pool_shape', None) != (self.img_H, self.img_W):
            self._pool = _PinnedPool((4, self.img_H, self.img_W))
            self._pool_shape = (self.img_H, self.img_W)
        slot = self._free_slots.pop()
        if self.engine.encode_frame(img, slot):
            self.engine.wait_frame_copied()
        self.memory = {start_frame_i: {'img': img, 'slot': slot,
                                       'result': FlowOUTrackingResult.identity((self.img_H, self.img_W), device=self.device)}}
        self.template_img = img.clone() if isinstance(img, torch.Tensor) else img.copy()      # (device-resident frames: bench / tapvid)
        meta = SimpleNamespace()
        meta.result = self.memory[start_frame_i]['result'].clone().cpu()
        return meta

    # ------------------------------------------------------------------------------------------
    def live_chains(self):
        """(delta, left_id) of the candidates for the current frame, in selection order
        [inf, ascending delta] (MFT.py:74-91 bookkeeping, :114 ordering)."""
        used, live = [], []
        for delta in self.C.deltas:
            left_id = self.current_frame_i - delta * self.time_direction
            if self.is_before_start(left_id):
                if np.isinf(delta):
                    left_id = self.start_frame_i
                else:
                    continue
            left_id = int(left_id)
            if left_id in used:
                continue
            used.append(left_id)
            live.append((delta, left_id))
        live.sort(key=lambda t: 0 if np.isinf(t[0]) else t[0])
        return live

    def track(self, input_img, debug=False, **kwargs):
        """input_img: (H,W,3) uint8 BGR (numpy, or a CUDA uint8 tensor already resident in HBM).
        meta.result: template -> current field, on CPU (kwarg device_result=True leaves it on the GPU
        and skips the device->host copy).  A frame in page-locked host memory is DMA'd in place; it may be
        refilled as soon as track() returns (like the reference, which copies at MFT/raft.py:45).
        Raises MftB200Error if a kernel of this or (device_result=True: of an earlier) frame aborted."""
        meta = SimpleNamespace()
        self.engine.error_flag_poll()                  # no sync: the mirror as of the last completed frame
        self.current_frame_i += self.time_direction
        right_id = self.current_frame_i
        H, W = self.img_H, self.img_W
        eng = self.engine

        slot = self._free_slots.pop()
        in_place = eng.encode_frame(input_img, slot)
        live = self.live_chains()
        K = len(live)
        right = torch.empty((K, 4, H, W), dtype=torch.float32, device=self.device)

        # flow cache protocol (MFT.py:189-230): finite deltas only unless C.cache_delta_infinity
        cache = self.flow_cache
        todo = []
        for k, (delta, left_id) in enumerate(live):
            use_cache = bool(np.isfinite(delta) or self.C.cache_delta_infinity)
            hit = False
            if use_cache and cache is not None:
                try:
                    f, o, s = cache.read(left_id, right_id)
                    assert f is not None
                    right[k, 0:2], right[k, 2:3], right[k, 3:4] = f.to(self.device), o.to(self.device), s.to(self.device)
                    hit = True
                except Exception:
                    hit = False
            if not hit:
                todo.append((k, left_id, use_cache))
        for i in range(0, len(todo), MAX_PAIRS):
            chunk = todo[i:i + MAX_PAIRS]
            lefts = [self.memory[left_id]['slot'] for _, left_id, _ in chunk]
            if len(chunk) == K:
                eng.refine(lefts, [slot] * len(chunk), out=right)
            else:
                res = eng.refine(lefts, [slot] * len(chunk))
                for j, (k, _, _) in enumerate(chunk):
                    right[k] = res[j]
        if cache is not None:
            for k, left_id, use_cache in todo:
                if use_cache:
                    cache.write(left_id, right_id, right[k, 0:2].clone(), right[k, 2:3].clone(), right[k, 3:4].clone())

        lefts = [self.memory[left_id]['result'].packed() for _, left_id in live]
        packed, index = chain_select(lefts, right, float(self.C.occlusion_threshold), want_index=bool(debug))
        result = FlowOUTrackingResult.from_packed(packed)
        eng.error_flag_async()                         # rides behind this frame's work, next to the result copy

        if kwargs.get('device_result', False):
            # the caller's own object AND storage, like the reference's clone (MFT.py:145): `.cpu()` or in-place edits
            # of meta.result must not reach the field stored in self.memory
            meta.result = result.clone()
            if in_place:
                eng.wait_frame_copied()                # the caller may refill its pinned frame buffer after this call
        else:
            host = self._pool.take()                  # pinned slot: asynchronous D2H, one stream sync, no extra copy
            if host is None:
                host = torch.empty((4, H, W), dtype=torch.float32)
            host.copy_(packed, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            eng.error_flag_poll()                      # this frame's flag: raises instead of returning garbage
            meta.result = FlowOUTrackingResult.from_packed(host)
        if debug:
            meta.selected_delta_i = index
            meta.used_deltas = [d for d, _ in live]

        self.memory[right_id] = {'img': input_img, 'result': result, 'slot': slot}
        self.cleanup_memory()
        return meta

    # ------------------------------------------------------------------------------------------
    def cleanup_memory(self):
        """Keep the template (if inf is a delta) and the last max-finite-delta frames (MFT.py:157-181)."""
        finite = [d for d in self.C.deltas if np.isfinite(d)]
        max_delta = max(finite) if finite else 0
        has_direct_flow = any(np.isinf(d) for d in self.C.deltas)
        for frame_i in list(self.memory.keys()):
            if frame_i == self.start_frame_i and has_direct_flow:
                continue
            if self.time_direction > 0 and frame_i + max_delta > self.current_frame_i:
                continue
            if self.time_direction < 0 and frame_i - max_delta < self.current_frame_i:
                continue
            self._free_slots.append(self.memory[frame_i]['slot'])
            del self.memory[frame_i]

    def is_before_start(self, frame_i):
        return ((self.time_direction > 0 and frame_i < self.start_frame_i) or
                (self.time_direction < 0 and frame_i > self.start_frame_i))
