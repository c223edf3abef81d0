"""Feature-slot exchange between ranks (run under torchrun, one rank per GPU): rank t % G encodes frame t, one in-place
all_gather per slot array and round; every rank's slots must equal, bit for bit, the slots of a second engine that
encodes every frame locally."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mft_b200 import engine as E  # noqa: E402
from mft_b200.synth import synthetic_video  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, G = dist.get_rank(), dist.get_world_size()
    H = W = int(os.environ.get('CHECK_SIZE', '256'))
    rounds = 4
    T = 1 + rounds * G
    frames = [torch.from_numpy(f).cuda() for f in synthetic_video(T, H, W, seed=7)]
    weights, _ = bench.load_weights()
    a, b = E.Engine(weights), E.Engine(weights)
    for eng in (a, b):
        eng.configure(H, W, max_pairs=2, n_slots=T, iters=12)
        eng.encode_frame(frames[0], 0)
    fa, fb = a.slot_tensors(), b.slot_tensors()
    names = ('fmap', 'net', 'inp')
    bad = 0
    for r in range(rounds):
        t0 = 1 + r * G
        for t in range(t0, t0 + G):
            b.encode_frame(frames[t], t)
        b.slot_tensors()
        a.encode_frame(frames[t0 + rank], t0 + rank)
        a.slot_tensors()
        for arr in fa:
            blk = arr[t0:t0 + G]
            dist.all_gather_into_tensor(blk.view(G * blk.shape[1], blk.shape[2]), arr[t0 + rank])
        torch.cuda.synchronize()
        for t in range(t0, t0 + G):
            for name, x, y in zip(names, fa, fb):
                if not torch.equal(x[t], y[t]):
                    d = (x[t].float() - y[t].float()).abs()
                    print(f'rank {rank}: round {r} slot {t} (encoded by rank {t - t0}) {name}: {int((d > 0).sum())} of {d.numel()} '
                          f'elements differ, max {d.max().item():.4g}', flush=True)
                    bad += 1
        # one refinement with the gathered features, like the sharded tracker: it may park / flush contexts
        out = a.refine([t0 + rank - 1], [t0 + rank])
        ref = b.refine([t0 + rank - 1], [t0 + rank])
        if not torch.equal(out, ref):
            print(f'rank {rank}: round {r} refine {t0 + rank - 1} -> {t0 + rank} differs, max {(out - ref).abs().max().item():.4g}', flush=True)
            bad += 1
    flag = torch.tensor([bad], device='cuda')
    dist.all_reduce(flag)
    if rank == 0:
        print('feature gather on', G, 'GPUs:', 'bit-identical to local encoding' if flag.item() == 0 else f'{int(flag.item())} MISMATCHES')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
