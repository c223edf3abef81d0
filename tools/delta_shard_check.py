"""Per-delta sharding inside a frame on real GPUs (run under torchrun, one rank per GPU):
every rank must end with fields bit-identical to the single-GPU tracker's, and the per-frame device time is reported
next to the single-GPU one (the refinement part is bounded by K / ceil(K / G))."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import engine as E  # noqa: E402
from mft_b200.dist import DeltaShardedTracker  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    size = int(os.environ.get('CHECK_SIZE', '256'))
    H = W = size
    T = 44
    deltas = [np.inf, 1, 2, 4, 8, 16, 32]
    frames = list(synthetic_video(T, H, W, seed=5))
    weights, _ = bench.load_weights()
    eng = E.Engine(weights)
    eng.configure(H, W, max_pairs=len(deltas), n_slots=T + 1, iters=12)
    eng.encode_frame(frames[0], 0)

    def flow_fn(t, live, out):
        eng.refine([left for _, left in live], [t] * len(live), out=out)      # straight into the gather buffer

    def select_fn(lefts, right):
        return E.chain_select(lefts, right, 0.02, want_index=False)[0]

    trk = DeltaShardedTracker(deltas, (H, W), flow_fn, select_fn, 'cuda')
    res, ms = {}, []
    for t in range(1, T):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.encode_frame(frames[t], t)               # every rank encodes every frame (slot = frame index here)
        res[t] = trk.track().clone()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    eng.check_device()
    ref = bench.make_tracker(weights)
    ref.C.deltas = deltas
    ref.init(frames[0])
    ok, ms1 = True, []
    for t in range(1, T):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        want = ref.track(frames[t], device_result=True).result.packed()
        b.record()
        torch.cuda.synchronize()
        ms1.append(a.elapsed_time(b))
        ok = ok and torch.equal(res[t], want)
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'delta-sharded tracking on {world} GPUs, {W}x{H}: {"bit-identical to single-GPU tracker" if flag.item() else "MISMATCH"}; '
              f'steady-state frame {np.mean(ms[-8:]):.3f} ms vs {np.mean(ms1[-8:]):.3f} ms on one GPU')
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == '__main__':
    main()
