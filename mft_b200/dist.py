"""Multi-GPU sharding of the tracking path (one process per GPU, torch.distributed).

The reference has no distributed code; every flow (left -> right) depends only on the
two frames, never on tracker state, so the path shards without touching its numerics:

* sequence sharding: independent sequences round-robin over ranks, no collective on
  the data path -- ``shard_items``; this is what bench.py scales (weak scaling).
* per-timestep flow sharding (offline video): rank t % G runs the batched K-pair
  refinement of frame t; ONE all_gather per round of G frames hands every rank the (K,4,H,W) blocks;
  the cheap sequential chain+select scan is replicated on all ranks so every replica of the tracker
  state stays identical -- ``FlowShardedTracker``.
* per-delta sharding inside one frame (online / low latency): chain k of frame t belongs to rank
  k % G (7 chains: 4+3 on two GPUs, 2+2+2+1 on four), every rank refines only its pairs, ONE all_gather of the
  (ceil(K/G),4,H,W) blocks per frame, then the fused chain+select runs replicated on all K fields, so every replica of
  the tracker state stays identical and the result is bit-identical to one GPU -- ``DeltaShardedTracker``.
  Latency bound K / ceil(K/G): 1.75x / 3.5x / 7x for the refinement part.
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


def chain_owner(k, world):
    return k % world


class DeltaShardedTracker:
    """Online tracking with the chains of ONE frame spread over the ranks (every rank sees every frame).

    flow_fn(t, live_subset) -> tensor (len(live_subset), 4, H, W): left->t fields of the given chains (on the GPU:
    Engine.refine on this rank's pairs; every rank has encoded frame t itself).  select_fn(lefts, right) -> (4,H,W):
    the fused chain+select over all K fields.  Same chain bookkeeping as FlowShardedTracker / MFT.track.
    """

    def __init__(self, deltas, shape, flow_fn, select_fn, device, max_chains=8, start=0):
        self.deltas, (self.H, self.W) = list(deltas), shape
        self.flow_fn, self.select_fn, self.device = flow_fn, select_fn, device
        self.start = start
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.per_rank = (max_chains + self.world - 1) // self.world
        self.results = {start: torch.zeros((4, self.H, self.W), dtype=torch.float32, device=device)}
        self.t = start

    live_chains = FlowShardedTracker.live_chains

    def track(self):
        """Advances one frame; returns its (4,H,W) result (identical on every rank)."""
        self.t += 1
        t, G = self.t, self.world
        live = self.live_chains(t)
        mine_idx = [k for k in range(len(live)) if chain_owner(k, G) == self.rank]
        mine = torch.zeros((self.per_rank, 4, self.H, self.W), dtype=torch.float32, device=self.device)
        if mine_idx:
            mine[:len(mine_idx)] = self.flow_fn(t, [live[k] for k in mine_idx])
        if G > 1:
            blocks = [torch.empty_like(mine) for _ in range(G)]
            dist.all_gather(blocks, mine)              # the single collective of the frame
        else:
            blocks = [mine]
        right = torch.stack([blocks[chain_owner(k, G)][k // G] for k in range(len(live))])
        lefts = [self.results[left] for _, left in live]
        self.results[t] = self.select_fn(lefts, right.contiguous())
        finite = [int(d) for d in self.deltas if not np.isinf(d)]
        old = t - (max(finite) if finite else 0)            # like MFT.cleanup_memory: template + the last max-delta frames
        if old != self.start:
            self.results.pop(old, None)
        return self.results[t]
