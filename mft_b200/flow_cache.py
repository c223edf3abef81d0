"""Device-resident flow cache.

The reference's FlowCache (MFT/utils/io.py:618-751) keeps (left_id, right_id) -> (flow, occlusion, sigma) in three
tiers (GPU -> RAM -> disk, the disk tier through a lossy 16-bit codec) and is consulted by the tracker for every finite
delta (MFT/MFT.py:189-230): the same frame pairs recur across the many start frames / directions of the TAP-Vid
evaluation (run_MFT_tapvid.py:164-181).  This one keeps the fields where the tracker consumes them -- in HBM, as ONE
packed (4,H,W) tensor per pair (what the fused chain+select kernel reads), fp32 (bit-exact replay) or fp16 (half the
footprint) -- behind the same read / write protocol, with least-recently-used eviction under a byte budget."""
import collections

import torch


class DeviceFlowCache:
    def __init__(self, max_bytes=8 << 30, dtype=torch.float32, device='cuda'):
        assert dtype in (torch.float32, torch.float16)
        self.max_bytes, self.dtype, self.device = int(max_bytes), dtype, device
        self.store = collections.OrderedDict()
        self.bytes = 0
        self.hits = self.misses = self.writes = self.evictions = 0

    def __len__(self):
        return len(self.store)

    def clear(self):
        self.store.clear()
        self.bytes = 0

    def read(self, left_id, right_id, **kwargs):
        """-> (flow (2,H,W), occlusion (1,H,W), sigma (1,H,W)) on the device, or (None, None, None) (io.py:655-672)."""
        key = (int(left_id), int(right_id))
        packed = self.store.get(key)
        if packed is None:
            self.misses += 1
            return None, None, None
        self.store.move_to_end(key)
        self.hits += 1
        p = packed if packed.dtype == torch.float32 else packed.float()
        return p[0:2], p[2:3], p[3:4]

    def write(self, left_id, right_id, flow, occlusion, sigma, **kwargs):
        """Stores a copy (io.py:674-698): the caller may reuse its tensors."""
        key = (int(left_id), int(right_id))
        packed = torch.cat([flow.reshape(2, *flow.shape[-2:]), occlusion.reshape(1, *flow.shape[-2:]),
                            sigma.reshape(1, *flow.shape[-2:])], 0).to(device=self.device)
        if self.dtype == torch.float16:
            # fp16 storage: flow resolution is 2^-10 relative (0.25 px at 256..512 px of motion, 0.03 px below 32 px),
            # sigma = sqrt(exp(u)) of a large predicted log-variance would overflow to inf: saturate at the fp16 maximum
            packed = packed.clamp(min=-65504.0, max=65504.0)
        packed = packed.to(dtype=self.dtype).contiguous()
        old = self.store.pop(key, None)
        if old is not None:
            self.bytes -= old.numel() * old.element_size()
        self.store[key] = packed
        self.bytes += packed.numel() * packed.element_size()
        self.writes += 1
        while self.bytes > self.max_bytes and len(self.store) > 1:
            _, ev = self.store.popitem(last=False)
            self.bytes -= ev.numel() * ev.element_size()
            self.evictions += 1
