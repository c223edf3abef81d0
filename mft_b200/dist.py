"""Multi-GPU sharding of the tracking path (one process per GPU, torch.distributed).

The reference has no distributed code; every flow (left -> right) depends only on the
two frames, never on tracker state, so the path shards without touching its numerics:

* sequence sharding: independent sequences round-robin over ranks, no collective on
  the data path -- ``shard_items``; this is what bench.py scales by default (weak scaling).
* per-timestep flow sharding (offline video): rank t % G runs the batched K-pair
  refinement of frame t straight into its block of ONE (G,K,4,H,W) buffer; ONE in-place
  ``all_gather_into_tensor`` per round of G frames fills the other blocks; the cheap sequential chain+select scan
  reads the blocks where the collective put them (no list gather, no stack / copy) and is replicated on all ranks, so
  every replica of the tracker state stays identical -- ``FlowShardedTracker``.  ``encode_fn`` optionally shards the
  per-frame encoders the same way (one feature all_gather per round, see bench.py --mode flow-shard).
* per-delta sharding inside one frame (online / low latency): the K chains of frame t are split into
  G contiguous ranges (7 chains: 4+3 on two GPUs, 2+2+2+1 on four), every rank refines only its pairs into its block of
  the (G*ceil(K/G),4,H,W) buffer, ONE in-place all_gather per frame, then the fused chain+select runs replicated on the
  first K fields of that buffer -- ``DeltaShardedTracker``.  Bit-identical to one GPU; latency bound K / ceil(K/G):
  1.75x / 3.5x / 7x for the refinement part.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_items(n_items, rank, world):
    """Indices of the items (sequences) owned by ``rank``: round-robin."""
    return list(range(rank, n_items, world))


def frame_owner(t, world):
    return t % world


def _gather_in_place(buf, rank, timer=None):
    """buf: (G, ...) contiguous; block ``rank`` holds this rank's contribution; fills the other blocks."""
    if timer is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    flat = buf.view((buf.shape[0] * buf.shape[1],) + tuple(buf.shape[2:]))      # concatenation form: block g = rows [g*K, (g+1)*K)
    if dist.get_backend() == 'nccl':
        dist.all_gather_into_tensor(flat, buf[rank])           # NCCL in-place form: the send block is block `rank` of the receive buffer
    else:
        dist.all_gather_into_tensor(flat, buf[rank].clone())   # (gloo in the CPU tests: separate send buffer)
    if timer is not None:
        b.record()
        timer.append((a, b))


def _call_flow(flow_fn, t, live, out):
    """flow_fn may write into ``out`` (GPU: Engine.refine(out=...)) or return a tensor (test stand-ins)."""
    res = flow_fn(t, live, out)
    if res is not None and res.data_ptr() != out.data_ptr():
        out.copy_(res)


class FlowShardedTracker:
    """Offline tracking of one video with per-timestep flow sharding.

    flow_fn(t, live, out): writes the left->t fields of the live chains of frame t into ``out`` (len(live),4,H,W)
    (on the GPU: Engine.refine(..., out=out); injected so the host logic is testable on CPU/gloo); may instead return
    them.  select_fn(lefts, right) -> (4,H,W): fused chain+select.  encode_fn(ts) (optional) is called once per round
    with the round's frame indices BEFORE any flow of the round (bench.py: encode the own frame, all_gather features).
    """

    def __init__(self, deltas, n_frames, shape, flow_fn, select_fn, device, max_chains=None, start=0, encode_fn=None,
                 time_gather=False):
        self.deltas, self.T, (self.H, self.W) = list(deltas), n_frames, shape
        self.flow_fn, self.select_fn, self.device, self.encode_fn = flow_fn, select_fn, device, encode_fn
        self.K = max_chains or len(self.deltas)
        self.start = start
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.gather_events = [] if time_gather else None

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

    def run(self, keep=None):
        """Returns {t: (4,H,W) result}; identical on every rank.  keep: frames whose result stays stored after they are
        no longer needed as a left result (None = all)."""
        return self.run_range(self.start + 1, self.start + self.T, keep)

    def run_range(self, t_begin, t_end, keep=None):
        """Tracks frames t_begin .. t_end - 1 (whole rounds: t_begin - start - 1 must be a multiple of the world size),
        continuing from the state left by earlier calls.  Returns the results dictionary (shared between calls)."""
        H, W, G = self.H, self.W, self.world
        assert (t_begin - self.start - 1) % G == 0, 'ranges start on a round boundary'
        if not hasattr(self, 'results'):
            self.results = {self.start: torch.zeros((4, H, W), dtype=torch.float32, device=self.device)}
            # ONE buffer for the whole run: rank g refines into block g, the in-place gather fills the other blocks
            self.buf = torch.zeros((G, self.K, 4, H, W), dtype=torch.float32, device=self.device)
        results, buf = self.results, self.buf
        finite = [int(d) for d in self.deltas if not np.isinf(d)]
        maxd = max(finite) if finite else 0
        t0 = t_begin
        while t0 < min(t_end, self.start + self.T):
            ts = [t for t in range(t0, min(t0 + G, self.start + self.T))]
            if self.encode_fn is not None:
                self.encode_fn(ts)
            my_t = next((t for t in ts if frame_owner(t - t0, G) == self.rank), None)
            if my_t is not None:
                live = self.live_chains(my_t)
                _call_flow(self.flow_fn, my_t, live, buf[self.rank, :len(live)])
            if G > 1:
                _gather_in_place(buf, self.rank, self.gather_events)      # the single collective of the round
            for i, t in enumerate(ts):                  # replicated sequential scan, straight out of the gather buffer
                live = self.live_chains(t)
                lefts = [results[left] for _, left in live]
                results[t] = self.select_fn(lefts, buf[i, :len(live)])
                old = t - maxd
                if keep is not None and old > self.start and old not in keep:
                    results.pop(old, None)
            t0 += G
        return results

    def gather_ms(self):
        """Per-round device time of the all_gather (after a synchronize), or [] when not timed."""
        return [a.elapsed_time(b) for a, b in (self.gather_events or [])]


def chain_range(rank, world, per_rank):
    return rank * per_rank, (rank + 1) * per_rank


class DeltaShardedTracker:
    """Online tracking with the chains of ONE frame spread over the ranks (every rank sees every frame).

    flow_fn(t, live_subset, out): the left->t fields of the given chains (on the GPU: Engine.refine on this rank's
    pairs; every rank has encoded frame t itself).  select_fn(lefts, right) -> (4,H,W): the fused chain+select over all
    K fields.  Same chain bookkeeping as FlowShardedTracker / MFT.track.
    """

    def __init__(self, deltas, shape, flow_fn, select_fn, device, max_chains=None, start=0, time_gather=False):
        self.deltas, (self.H, self.W) = list(deltas), shape
        self.flow_fn, self.select_fn, self.device = flow_fn, select_fn, device
        self.start = start
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.per_rank = ((max_chains or len(self.deltas)) + self.world - 1) // self.world
        self.results = {start: torch.zeros((4, self.H, self.W), dtype=torch.float32, device=device)}
        # rank g's chains are rows [g * per_rank, (g + 1) * per_rank): after the gather the buffer IS the (K,4,H,W) operand
        self.buf = torch.zeros((self.world, self.per_rank, 4, self.H, self.W), dtype=torch.float32, device=device)
        self.t = start
        self.gather_events = [] if time_gather else None

    live_chains = FlowShardedTracker.live_chains
    gather_ms = FlowShardedTracker.gather_ms

    def track(self):
        """Advances one frame; returns its (4,H,W) result (identical on every rank)."""
        self.t += 1
        t, G = self.t, self.world
        live = self.live_chains(t)
        lo, hi = chain_range(self.rank, G, self.per_rank)
        mine = live[lo:hi]
        if mine:
            _call_flow(self.flow_fn, t, mine, self.buf[self.rank, :len(mine)])
        if G > 1:
            _gather_in_place(self.buf, self.rank, self.gather_events)     # the single collective of the frame
        right = self.buf.view(G * self.per_rank, 4, self.H, self.W)[:len(live)]
        lefts = [self.results[left] for _, left in live]
        self.results[t] = self.select_fn(lefts, right)
        finite = [int(d) for d in self.deltas if not np.isinf(d)]
        old = t - (max(finite) if finite else 0)            # like MFT.cleanup_memory: template + the last max-delta frames
        if old != self.start:
            self.results.pop(old, None)
        return self.results[t]
