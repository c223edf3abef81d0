"""Per-timestep flow sharding on real GPUs (run under torchrun, one rank per GPU):
every rank must end with fields bit-identical to the single-GPU tracker's (same kernels, same inputs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import engine as E  # noqa: E402
from mft_b200.dist import FlowShardedTracker  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W, T = 128, 192, 14
    deltas = [np.inf, 1, 2, 4, 8]
    frames = list(synthetic_video(T, H, W, seed=5))
    weights, _ = bench.load_weights()
    eng = E.Engine(weights)
    eng.configure(H, W, max_pairs=len(deltas), n_slots=T + 1, iters=12)
    for i, f in enumerate(frames):          # every rank encodes the (cheap) per-frame features it may need
        eng.encode_frame(f, i)

    def flow_fn(t, live, out):
        eng.refine([left for _, left in live], [t] * len(live), out=out)      # straight into the gather buffer

    def select_fn(lefts, right):
        return E.chain_select(lefts, right, 0.02, want_index=False)[0]

    res = FlowShardedTracker(deltas, T, (H, W), flow_fn, select_fn, 'cuda').run()
    eng.check_device()
    # reference: the plain single-GPU tracker on this rank
    trk = bench.make_tracker(weights)
    trk.C.deltas = deltas
    trk.init(frames[0])
    ok = True
    for t in range(1, T):
        want = trk.track(frames[t], device_result=True).result.packed()
        ok = ok and torch.equal(res[t], want)
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'flow-sharded tracking on {world} GPUs: {"bit-identical to single-GPU tracker" if flag.item() else "MISMATCH"}')
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == '__main__':
    main()
