"""Multi-GPU sharding of the tracking path (one process per GPU, torch.distributed).

The reference has no distributed code; every flow (left -> right) depends only on the
two frames, never on tracker state, so the path shards without touching its numerics:

* sequence sharding: independent sequences round-robin over ranks, no collective on
  the data path -- ``shard_items``; this is what bench.py scales (weak scaling).
* per-timestep flow sharding (offline video): rank t % G runs the batched K-pair
  refinement of frame t; ONE all_gather per round of G frames hands every rank the (K,4,H,W) blocks;
  the cheap sequential chain+select scan is replicated on all ranks so every replica of the tracker
  state stays identical -- ``FlowShardedTracker``.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_items(n_items, rank, world):
    """Indices of the items (sequences) owned by ``rank``: round-robin."""
    return list(range(rank, n_items, world))


def frame_owner(t, world):
    return t % world


class FlowShardedTracker:
    """Offline tracking of one video with per-timestep flow sharding.

    flow_fn(t, live) -> tensor (len(live), 4, H, W): left->t fields for the live chains of frame t
    (on the GPU: encode + Engine.refine; injected so the host logic is testable on CPU/gloo).
    select_fn(lefts, right) -> (4,H,W): fused chain+select.
    """

    def __init__(self, deltas, n_frames, shape, flow_fn, select_fn, device, max_chains=8, start=0):
        self.deltas, self.T, (self.H, self.W) = list(deltas), n_frames, shape
        self.flow_fn, self.select_fn, self.device = flow_fn, select_fn, device
        self.K = max_chains
        self.start = start
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0

    def live_chains(self, t):
        used, live = [], []
        for d in self.deltas:
            left = self.start if np.isinf(d) else t - int(d)
            if not np.isinf(d) and left < self.start:
                continue
            if left in used:
                continue
            used.append(left)
            live.append((d, int(left)))
        live.sort(key=lambda x: 0 if np.isinf(x[0]) else x[0])
        return live

    def run(self):
        """Returns {t: (4,H,W) result}; identical on every rank."""
        H, W, G = self.H, self.W, self.world
        results = {self.start: torch.zeros((4, H, W), dtype=torch.float32, device=self.device)}
        t0 = self.start + 1
        while t0 < self.start + self.T:
            ts = [t for t in range(t0, min(t0 + G, self.start + self.T))]
            mine = torch.zeros((self.K, 4, H, W), dtype=torch.float32, device=self.device)
            my_t = next((t for t in ts if frame_owner(t - t0, G) == self.rank), None)
            if my_t is not None:
                live = self.live_chains(my_t)
                mine[:len(live)] = self.flow_fn(my_t, live)
            if G > 1:
                blocks = [torch.empty_like(mine) for _ in range(G)]
                dist.all_gather(blocks, mine)          # the single collective of the path
            else:
                blocks = [mine]
            for i, t in enumerate(ts):                  # replicated sequential scan
                live = self.live_chains(t)
                lefts = [results[left] for _, left in live]
                results[t] = self.select_fn(lefts, blocks[i][:len(live)].contiguous())
            t0 += G
        return results
