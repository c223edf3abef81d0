#!/usr/bin/env python
"""Headline benchmark: dense-track frames/sec at 512x512 with the 7 delta chains (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = MFT.track() of one steady-state frame (t > 32, K = 7 live chains): encode the new frame
once, one batched 7-pair RAFT-OU refinement (12 iterations), fused chain + select.  One rank per GPU
(torchrun), each rank tracks its own synthetic sequence (sequence sharding):
weak scaling, no data-path collective; timing = max over ranks of the device time.

value   frames/s, inputs resident in HBM: ONE CUDA-event bracket around the K steps, L2 flushed between steps (160 MiB write, inside the
        bracket), the context encoder that trails each frame on a side stream joined before the closing event
e2e     frames/s through the public API (mft_b200.MFT.MFT.track) with HOST numpy frames: pinned H2D of
        the frame and D2H of the (4,H,W) result inside the timed region
roofline  tensor-core conv kernel family: algorithmic FLOPs (minimal formulation) over the
        summed per-launch event time of the conv launches of one step
cpu_baseline  the reference's CPU path timed on this box's host cores: the UNMODIFIED reference tracker (through
        oracle/ref_bridge.py's device-agnostic subclass) when a reference checkout is reachable ($MFT_REFERENCE_ROOT,
        /root/reference, baseline/_ref) -> kind "reference"; otherwise the oracle port -> kind "port"
--impl reference  the same CPU path as its own arm
parity  one checked frame of THIS run: frame 33 of the benchmark video (first steady-state frame) against the vector
        recorded from the unmodified reference (tests/golden/track_synth_512.npz)

Other workloads of BASELINE.json (not run by the driver's default command line):
  --mode flow-shard   config 4: ONE synthetic 1024x1024 video, 32 GRU iterations, per-timestep flows sharded over the
                      ranks, one in-place all_gather of the (7,4,H,W) blocks per round, replicated scan; strong scaling
  --mode tapvid       config 3: synthetic TAP-Vid-schema dataset (256x256 -> 512x512), first + strided query modes,
                      forward + backward tracking with the device flow cache, sequences round-robin over ranks
  --size HxW          config 5: e.g. --size 1080x1920 (one sequence per GPU, full delta set)
  --mode demo         config 2: the reference's demo video, full length, 512x512, full delta set, one GPU: init + track of
                      every frame as demo.py does (numpy frames in, CPU result out), wall clock and device time
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DELTAS = [np.inf, 1, 2, 4, 8, 16, 32]
STEADY = 33            # frames tracked before the timed region so that all 7 chains are live


def flops_per_frame(H, W, iters=12, K=7):
    """Minimal algorithmic FLOPs of one steady-state frame."""
    H, W = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    n = (H // 8) * (W // 8)
    enc = (H // 2) * (W // 2) * (18816 + 4 * 73728) + (H // 4) * (W // 4) * 620544 + n * (1130496 + 65536)
    per_pair = 2 * n * n * 256 + (iters - 1) * n * 5351936 + n * 6236672 + n * 3287808
    return 2 * enc + K * per_pair


def load_weights():
    """The shipped checkpoint when it is reachable, else a random initialisation of the architecture -- through the
    product's own loader (mft_b200/weights.py); nothing on the GPU arm touches oracle/."""
    from mft_b200 import weights as WT
    path = WT.find_checkpoint()
    if path is not None:
        return WT.load_checkpoint(path), 'shipped checkpoint'
    return WT.random_init(0), 'random init of the architecture (checkpoint not reachable)'


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU, sampled DURING the timed regions: NVML polled in-process
    every ~10 ms (a fresh `nvidia-smi` needs longer to start than a 20-step timed region lasts); `nvidia-smi -lms` only
    as the fallback when NVML cannot be loaded."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, gpu_index):
        self.idx, self.proc, self.nvml, self.handle = gpu_index, None, None, None
        self.sm, self.mx, self.reasons = [], [], set()
        self.active, self.quit, self.thread, self.source = False, False, None, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:                                                # CUDA_VISIBLE_DEVICES may renumber: go by UUID
            import torch
            uuid = 'GPU-' + str(torch.cuda.get_device_properties(self.idx).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        return pynvml, h

    def start(self):
        """Begin polling (idle until resume())."""
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.source = 'nvml'
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.idx), '-lms', '20'], stdout=subprocess.PIPE, text=True)
            self.source = 'nvidia-smi'
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def resume(self):
        self.active = True

    def pause(self):
        self.active = False

    def _poll_nvml(self):
        nv, h = self.nvml, self.handle
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, 'hw_slowdown'),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, 'hw_thermal_slowdown'),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, 'sw_thermal_slowdown'),
                (nv.nvmlClocksEventReasonSwPowerCap, 'sw_power_cap')]
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self.quit:
            if self.active:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    if mx is not None:
                        self.mx.append(mx)
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in bits:
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.01)

    def _read_smi(self):
        for line in self.proc.stdout:
            if not self.active:
                continue
            r = [x.strip() for x in line.split(',')]
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                for n, v in zip(self.NAMES, r[4:8]):
                    if v.lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass

    def stop(self):
        if self.source is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml and nvidia-smi unavailable'], 'samples': 0}
        self.quit = True
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
        busy = [s for s in self.sm if s > 500] or self.sm
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': self.source,
                'sampled': 'during the device-timed steps and the end-to-end steps'}


def make_tracker(weights, iters=12, deltas=None):
    from mft_b200.config import Config
    from mft_b200.MFT import MFT
    from mft_b200.raft import RAFTWrapper
    fc = Config()
    fc.of_class = RAFTWrapper
    fc.model = weights
    fc.flow_iters = iters
    fc.raft_params = {'occlusion_module': 'separate_with_uncertainty', 'small': False, 'mixed_precision': False}
    C = Config()
    C.tracker_class = MFT
    C.flow_config = fc
    C.deltas = list(deltas if deltas is not None else DELTAS)
    C.occlusion_threshold = 0.02
    return MFT(C)


def workload_name(H, W, iters=12):
    return f'synthetic {W}x{H} video, deltas [inf,1,2,4,8,16,32], {iters} GRU iters, steady state (7 live chains)'


# ------------------------------------------------------------------------------------------------
def cpu_reference_arm(H, W, steps, warmup, threads, budget_s=240.0):
    """Steady-state frames through the reference's CPU path: 7 RAFT-OU forwards (3 encoder passes each, as the
    reference does) + 7 chains + selection per step.  With a reference checkout reachable this is the UNMODIFIED
    reference tracker (MFT/MFT.py:55-154 around MFT/RAFT/core/raft.py, device string 'cpu' through
    oracle/ref_bridge.py); otherwise the oracle port.  The tracker memory is pre-filled with identity fields for
    frames 1..32 so that every timed step is a t > 32 frame (7 live chains) without first tracking 33 frames on the
    CPU -- the cost of a frame does not depend on the stored fields' values.
    Returns (seconds per frame, chains timed per step, kind)."""
    import torch
    from mft_b200.synth import synthetic_video
    torch.set_num_threads(threads)
    n = STEADY + warmup + steps
    frames = list(synthetic_video(n + 1, H, W, seed=1234))
    kind = 'port'
    try:
        from oracle import ref_bridge as R
        if R.available():
            kind = 'reference'
    except Exception:
        kind = 'port'
    if kind == 'reference':
        import warnings
        warnings.filterwarnings('ignore')
        model = R.build_reference_model()
        trk = R.build_reference_tracker(model, DELTAS)
        trk.init(frames[0])
        ident = trk.memory[0]['result']
        for t in range(1, STEADY):
            trk.memory[t] = {'img': frames[t], 'result': ident.clone()}
        trk.current_frame_i = STEADY - 1

        def step(frame):
            trk.track(frame)
            return 7 if len(trk.C.deltas) == 7 else len(trk.C.deltas)

        def shrink(pairs):
            trk.C.deltas = DELTAS[:pairs]
    else:
        from oracle import mft_oracle as O
        from mft_b200 import weights as WT
        path = WT.find_checkpoint()
        W_ = O.load_checkpoint(path) if path else O.seeded_weights(0)
        trk = O.OracleTracker(W_, deltas=DELTAS, fast_lookup=True)
        trk.init(frames[0])
        zero = trk.memory[0]['result']
        for t in range(1, STEADY):
            trk.memory[t] = dict(img=frames[t], result=zero)
        trk.cur = STEADY - 1

        def step(frame):
            return len(trk.track(frame).live)

        def shrink(pairs):
            trk.deltas = DELTAS[:pairs]
    times, pairs = [], 7
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        live = step(frames[STEADY + i])
        assert live == pairs, (live, pairs)
        times.append((time.perf_counter() - t0) * 7.0 / pairs)
        if i == 0 and times[0] * (warmup + steps) > budget_s:
            # bounded sample: keep the run inside the budget by tracking fewer chains per step and
            # scaling the step time to the full 7 chains (cost is linear in the number of pairs)
            pairs = int(max(1, min(7, budget_s / (times[0] / 7.0 * (warmup + steps)))))
            shrink(pairs)
    return float(np.mean(times[warmup:])), pairs, kind


def parse_size(s, default):
    if not s:
        return default, default
    if 'x' in str(s):
        h, w = str(s).lower().split('x')
        return int(h), int(w)
    return int(s), int(s)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    H, W = parse_size(args.size, 512)
    threads = os.cpu_count()
    sec, pairs, kind = cpu_reference_arm(H, W, args.steps, args.warmup, threads)
    fps = 1.0 / sec
    what = ('the unmodified reference tracker on CPU (MFT.MFT.track around the reference RAFT module, oracle/ref_bridge.py)'
            if kind == 'reference' else 'oracle port of the reference PyTorch path (no reference checkout on this box)')
    line = {
        'impl': 'reference', 'metric': 'dense-track frames/sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(H, W), 'device': 'host CPU'},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': kind,
                         'sample': f'{args.steps} steady-state frames ({pairs} of 7 RAFT-OU forwards + chain + select '
                                   f'timed per step, scaled to 7; tracker memory pre-filled with identity fields), {what}, torch CPU fp32'},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return world, rank, local


def make_barrier(world):
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return barrier


def peaks_info():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    src = 'MEASURED_PEAKS.json bf16_tflops_sustained (fp16 runs at the bf16 rate)' if peaks else 'fallback 1400 TFLOP/s sustained'
    return peak_tf, src


def ncu_traffic():
    """DRAM bytes per conv_prog_kernel launch from the committed `ncu --set full` capture of this build (a profiler cannot
    run inside the timed process); None when no capture of this round is present."""
    for name in ('r2_ncu_conv_prog.json', 'r1_ncu_conv_prog.json'):
        try:
            d = json.load(open(os.path.join(ROOT, 'profiles', name)))
            return d['dram_bytes_per_launch'], f'profiles/{name}'
        except Exception:
            continue
    return None, None


def parity_check(result_packed, index, frames):
    """One checked frame of this run (frame 33 of the seed-1234 512x512 benchmark video) against the vector recorded
    from the unmodified reference (committed fixture; no oracle code involved)."""
    import zlib
    try:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'track_synth_512.npz'))
    except Exception as ex:
        return {'checked': False, 'why': f'fixture not readable: {ex}'}
    crc = [zlib.crc32(np.ascontiguousarray(f).tobytes()) & 0xffffffff for f in frames[:34]]
    if crc != [int(c) for c in g['frame_crc'][:34]]:
        return {'checked': False, 'why': 'the synthetic frames generated on this box differ from the recorded ones (cv2 build)'}
    got = result_packed.cpu().numpy()
    ref = g['result_33']
    epe = np.sqrt(((got[:2, ::2, ::2] - ref[:2]) ** 2).sum(0))
    return {'checked': True, 'frame': 33, 'against': 'unmodified reference on CPU fp32 (tests/golden/track_synth_512.npz)',
            'epe_mean_px': float(epe.mean()), 'epe_median_px': float(np.median(epe)), 'epe_p995_px': float(np.quantile(epe, 0.995)),
            'occlusion_mean_abs': float(np.abs(got[2, ::2, ::2] - ref[2]).mean()),
            'index_agreement': float((index.cpu().numpy() == g['index_33']).mean()),
            'flow_median_magnitude_px': float(g['stats'][32][4])}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mft_b200.synth import synthetic_video

    world, rank, local = dist_setup()
    H, W = parse_size(args.size, 512)
    K, Wm = args.steps, args.warmup
    weights, wsrc = load_weights()
    for kv in filter(None, os.environ.get('BENCH_GLOBAL_OPTIONS', '').split(',')):      # A/B runs of configure-time knobs, e.g. prog_split_n=1
        from mft_b200 import engine as _E
        k, v = kv.split('=')
        _E.set_global_option(k.strip(), int(v))
    tracker = make_tracker(weights)
    n_frames = 1 + STEADY + 4 * (Wm + K) + 6              # room for one repeated measurement
    frames = list(synthetic_video(n_frames, H, W, seed=1234 + rank))
    dev_frames = [torch.from_numpy(f).cuda() for f in frames]
    # end-to-end arm: the caller's frames live in page-locked host memory (numpy views of pinned tensors), so the
    # per-frame host->device copy inside the timed region is one asynchronous DMA from the caller's buffer
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    host_frames = [p.numpy() for p in pinned]
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device='cuda')      # > the 126 MB L2
    barrier = make_barrier(world)

    # ---- reach the steady state -------------------------------------------------------------------
    tracker.init(frames[0])
    eng = tracker.engine
    for kv in filter(None, os.environ.get('BENCH_ENGINE_OPTIONS', '').split(',')):      # A/B runs, e.g. defer_context=0
        k, v = kv.split('=')
        eng.set_option(k.strip(), int(v))
    t = 1
    parity = None
    for i in range(STEADY):
        last = i == STEADY - 1
        meta = tracker.track(dev_frames[t], device_result=True, debug=last)
        if last and rank == 0 and (H, W) == (512, 512):
            parity = parity_check(meta.result.packed(), meta.selected_delta_i, frames)
        t += 1
    eng.check_device()

    # ---- the two timed regions; repeated ONCE if the clocks sampled during them report a slow-down -------------
    BAD = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    attempts = []
    for attempt in range(2):
        # value: device-resident inputs; ONE event pair brackets the K steps INCLUDING the L2 flush between them and the
        # context encoder that trails each frame on the engine's side stream (joined before the closing event): every
        # kernel of the K frames lies inside the bracket, and frame t+1 overlaps cnet(t) as it does in production
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(Wm):
            tracker.track(dev_frames[t], device_result=True)
            t += 1
        barrier()
        sampler.resume()
        launches0 = eng.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            flush.zero_()
            tracker.track(dev_frames[t], device_result=True)
            t += 1
        eng.slot_tensors()                         # orders the stream behind the trailing context encoder of the last frame
        b.record()
        barrier()
        sampler.pause()
        launches = eng.launch_count() - launches0 + K          # + one chain_select launch per step
        dev_ms = a.elapsed_time(b)
        # e2e: host frames through the public API, wall clock
        for _ in range(Wm):
            tracker.track(host_frames[t])
            t += 1
        barrier()
        sampler.resume()
        t0 = time.perf_counter()
        per_frame = []
        for _ in range(K):
            t1 = time.perf_counter()
            meta = tracker.track(host_frames[t])
            per_frame.append(time.perf_counter() - t1)
            t += 1
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 else None
        redo = torch.tensor([1 if (clocks and BAD & set(clocks['reasons'])) else 0], device='cuda')
        if world > 1:
            dist.broadcast(redo, 0)                                 # every rank repeats, or none
        attempts.append(clocks)
        if not int(redo.item()) or attempt == 1:
            break
        time.sleep(5.0)
    if clocks is not None and len(attempts) > 1:
        clocks['rejected_first_attempt'] = attempts[0]
    if os.environ.get('BENCH_DEBUG'):
        print('e2e per-frame ms:', [round(x * 1e3, 2) for x in per_frame], file=sys.stderr)
    assert tuple(meta.result.flow.shape) == (2, H, W) and not meta.result.flow.is_cuda
    eng.check_device()
    # ---- live per-launch profile of one step (rank 0) ----------------------------------------------
    conv_ms = other_ms = 0.0
    conv_n = 0
    zr_ms, zr_n = 0.0, 0
    by_tag = {}
    if rank == 0:
        eng.set_option('profile', 1)
        reps = 3
        for _ in range(reps):
            tracker.track(dev_frames[t], device_result=True)
            t += 1
        for ms, kind, layer in eng.profile_steps():
            if layer == 200:             # conv_prog_kernel: the persistent launch of one GRU iteration's 11 convolutions
                zr_ms += ms
                zr_n += 1
            by_tag.setdefault((kind, layer), []).append(ms)
        (conv_ms, other_ms), (conv_n, _) = eng.profile_fetch()
        conv_ms, other_ms, conv_n = conv_ms / reps, other_ms / reps, conv_n // reps
        eng.set_option('profile', 0)

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_src = peaks_info()
    F = flops_per_frame(H, W)
    family = F / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
    # dominant kernel: conv_prog_kernel, the persistent tcgen05 launch that runs the 11 convolutions of one GRU iteration
    # (motion encoder 5, SepConvGRU 4 fused z|r / q, flow head 2) for all 7 pairs with tile-level dataflow: 12 launches
    # per step, 5 351 936 algorithmic FLOPs per coarse pixel each
    npx = ((H + 7) // 8) * ((W + 7) // 8)
    zr_flops = 7.0 * npx * 5351936
    achieved = zr_flops / (zr_ms / zr_n * 1e-3) / 1e12 if zr_n else None
    traffic, traffic_src = ncu_traffic()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count()
        sec, pairs, kind = cpu_reference_arm(H, W, 2, 1, threads, budget_s=60.0)
        cpu = {'value': 1.0 / sec, 'unit': 'frames/s', 'cores': threads, 'kind': kind,
               'sample': f'2 steady-state frames after 1 warm-up ({pairs} of 7 RAFT-OU forwards + chain + select per step, scaled to 7; '
                         'tracker memory pre-filled with identity fields), ' +
                         ('the unmodified reference tracker on CPU' if kind == 'reference' else 'oracle port of the reference PyTorch path') +
                         ', torch CPU fp32'}

    line = {
        'metric': 'dense-track frames/sec', 'value': world * K / (dev_ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world,
        'steps': K, 'warmup': Wm, 'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
        'config': {'workload': workload_name(H, W),
                   'sharding': 'one independent sequence per GPU', 'weights': wsrc,
                   'cache': '160 MiB L2 flush (L2 = 126 MB) between timed steps, INSIDE the one event bracket around the K steps',
                   'e2e': 'MFT.track(frame) with uint8 frames in pinned host memory, result returned as CPU tensors',
                   'arithmetic': 'fp16 tensor-core operands, fp32 accumulate / recurrent state / outputs'},
        'e2e': {'value': world * K / e2e_s, 'unit': 'frames/s', 'h2d_bytes_per_step': H * W * 3, 'd2h_bytes_per_step': 16 * H * W},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                     'frac': (achieved / peak_tf) if achieved else None, 'traffic': traffic, 'traffic_source': traffic_src,
                     'kernel': 'conv_prog_kernel (persistent tcgen05 implicit-GEMM program: the 11 convolutions of one GRU '
                               f'iteration, M={7 * npx} pixel rows, tile-level dataflow between layers)',
                     'peak_source': peak_src, 'flops_per_launch': zr_flops, 'launches_per_step': zr_n // 3 if zr_n else 0,
                     'us_per_launch': (zr_ms / zr_n * 1e3) if zr_n else None,
                     'conv_family': {'achieved': family, 'frac': (family / peak_tf) if family else None,
                                     'flops_per_step': F, 'launches_per_step': conv_n, 'ms_per_step': conv_ms,
                                     'note': 'all tcgen05 conv/GEMM launches of one step, minimal algorithmic FLOPs (BASELINE.md)'},
                     'other_kernels_ms_per_step': other_ms},
        'clocks': clocks,
    }
    if parity is not None:
        line['parity'] = parity
    if cpu is not None:
        line['cpu_baseline'] = cpu
    print(json.dumps(line), flush=True)
    if os.environ.get('BENCH_DETAIL'):
        from mft_b200 import weights as _w
        for (kind, layer), v in sorted(by_tag.items(), key=lambda kv: -sum(kv[1])):
            name = _w.LAYER_NAMES[layer] if 0 <= layer < len(_w.LAYER_NAMES) else {100: 'corr GEMM', 200: 'iteration program', 201: 'full program', 202: 'heads program'}.get(layer, 'bandwidth kernel')
            print(f'  {name:24s} kind {kind}  n {len(v) // 3:3d}/step  {sum(v) / 3 * 1e3:9.1f} us/step  {np.mean(v) * 1e3:8.1f} us each', file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_flow_shard(args):
    """BASELINE config 4 / SURVEY 8e(ii): ONE long video, frame t's batched 7-pair refinement on rank t % G, one in-place
    all_gather of the (7,4,H,W) blocks per round of G frames, replicated chain+select scan.  The per-frame encoders are
    sharded the same way (rank t % G encodes frame t, one feature all_gather per round).  Strong scaling: the K timed
    frames are the same whatever G."""
    import torch
    import torch.distributed as dist
    from mft_b200 import engine as E
    from mft_b200.dist import FlowShardedTracker
    from mft_b200.synth import synthetic_video

    world, rank, local = dist_setup()
    G = world
    H, W = parse_size(args.size, 1024)
    iters = args.iters or 32
    K, Wm = args.steps, args.warmup
    K = (K + G - 1) // G * G                                   # whole rounds
    pre = (STEADY - 1 + Wm + G - 1) // G * G                   # frames 1..pre build the state (>= 32: all 7 chains live afterwards)
    T = 1 + pre + K
    weights, wsrc = load_weights()
    frames = list(synthetic_video(T, H, W, seed=1234))
    n_slots = 1 + (T - 1 + G - 1) // G * G
    eng = E.Engine(weights)
    eng.configure(H, W, max_pairs=7, n_slots=n_slots, iters=iters)
    barrier = make_barrier(world)
    local_encode = bool(os.environ.get('BENCH_FLOWSHARD_LOCAL_ENCODE'))      # bisecting aid: every rank encodes every frame itself
    own = {t: torch.from_numpy(frames[t]).cuda() for t in range(T) if t == 0 or (t - 1) % G == rank or local_encode}
    eng.encode_frame(own[0], 0)                                # the template: every rank encodes it itself
    feats = eng.slot_tensors()
    feat_events = []

    def encode_fn(ts):
        t0 = ts[0]
        mine = t0 + rank
        if local_encode:
            for t in ts:
                eng.encode_frame(own[t], t)
            return
        if mine in own:
            eng.encode_frame(own[mine], mine)                  # slot index == frame index
        if G > 1:
            eng.slot_tensors()                                 # stream ordered behind the trailing context encoder
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for arr in feats:                                  # fmap, net, inp: in-place all_gather of the round's G slots
                blk = arr[t0:t0 + G]
                dist.all_gather_into_tensor(blk.view(G * blk.shape[1], blk.shape[2]), arr[t0 + rank])
            b.record()
            feat_events.append((a, b))

    def flow_fn(t, live, out):
        eng.refine([left for _, left in live], [t] * len(live), out=out)      # straight into the gather buffer

    def select_fn(lefts, right):
        return E.chain_select(lefts, right, 0.02, want_index=False)[0]

    def tracker(n, timed):
        return FlowShardedTracker(DELTAS, n, (H, W), flow_fn, select_fn, 'cuda', encode_fn=encode_fn, time_gather=timed)

    # state-building part (frames 1..pre), then the K timed frames continue the same scan
    trk = tracker(T, True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    results = trk.run_range(1, 1 + pre)
    barrier()
    n_gather0, n_feat0 = len(trk.gather_events), len(feat_events)
    sampler.resume()
    launches0 = eng.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    results = trk.run_range(1 + pre, T)
    b.record()
    barrier()
    sampler.pause()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = a.elapsed_time(b)
    launches = eng.launch_count() - launches0 + K
    eng.check_device()
    gather_ms = trk.gather_ms()[n_gather0:]
    feat_ms = [x.elapsed_time(y) for x, y in feat_events[n_feat0:]]
    # ---- in-run identity check: the plain single-GPU tracker over the same video, on every rank -------------------------------
    ident = None
    if not args.no_identity_check:
        ref = make_tracker(weights, iters=iters)
        ref.init(frames[0])
        ok, first_bad, worst = True, -1, 0.0
        for t in range(1, T):
            want = ref.track(frames[t], device_result=True).result.packed()
            if not torch.equal(results[t], want):
                if first_bad < 0:
                    first_bad = t
                d = (results[t] - want).abs()
                worst = max(worst, float(d[torch.isfinite(d)].max().item()) if torch.isfinite(d).any() else float('inf'))
                if t > pre:
                    ok = False
        if first_bad >= 0:
            print(f'[flow-shard identity] rank {rank}: first differing frame {first_bad}, largest |difference| {worst:.6g}', file=sys.stderr)
        flag = torch.tensor([1 if ok else 0], device='cuda')
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ident = bool(flag.item())
    if world > 1:
        tt = torch.tensor([dev_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = tt.item()
    if rank == 0:
        peak_tf, peak_src = peaks_info()
        F = flops_per_frame(H, W, iters=iters)
        ach = F * K / (dev_ms * 1e-3) / 1e12 / world
        block_bytes = 7 * 4 * H * W * 4
        line = {
            'metric': 'dense-track frames/sec', 'value': K / (dev_ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world, 'steps': K,
            'warmup': pre, 'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': f'ONE synthetic {W}x{H} video, deltas [inf,1,2,4,8,16,32], {iters} GRU iters, steady state (7 live chains), '
                                   f'{K} timed frames after {pre} state-building frames',
                       'sharding': 'per-timestep flows: frame t on rank t % G; one in-place all_gather_into_tensor of the (7,4,H,W) blocks per '
                                   'round of G frames + one feature all_gather (fmap / net / inp slots); chain+select scan replicated',
                       'weights': wsrc, 'cache': 'per-step working set (7 correlation pyramids = 1.2 GB at 1024^2) exceeds the 126 MB L2; no flush'},
            'gpu_launches': int(launches),
            'collective': {'op': 'all_gather_into_tensor (NCCL, in place)', 'bytes_per_rank_per_round': block_bytes,
                           'rounds': len(gather_ms), 'flow_gather_us_per_round': float(np.mean(gather_ms) * 1e3) if gather_ms else 0.0,
                           'feature_gather_us_per_round': float(np.mean(feat_ms) * 1e3) if feat_ms else 0.0},
            'identical_to_single_gpu': ident,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                         'traffic': None, 'kernel': 'whole frame per GPU (minimal algorithmic FLOPs of the K frames / time / GPUs)',
                         'peak_source': peak_src},
            'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_tapvid(args):
    """BASELINE config 3: TAP-Vid style evaluation (synthetic dataset with the TAP-Vid schema, 256x256 -> 512x512), query modes
    first + strided, forward + backward tracking with the device flow cache, sequences round-robin over the ranks, no
    data-path collective."""
    import torch
    import torch.distributed as dist
    from mft_b200 import tapvid as TV
    from mft_b200.dist import shard_items
    from mft_b200.flow_cache import DeviceFlowCache

    world, rank, local = dist_setup()
    H, W = parse_size(args.size, 512)
    n_seq, n_frames = max(world, args.sequences), args.frames
    data = TV.synthetic_dataset(n_seq, n_frames, 32, 256, seed=1234)
    names = sorted(data)
    weights, wsrc = load_weights()
    tracker = make_tracker(weights)
    barrier = make_barrier(world)
    mine = [names[i] for i in shard_items(len(names), rank, world)]

    def prepare(name):
        d = data[name]
        video = TV.resize_video(TV.resize_video(d['video'], (256, 256)), (H, W))            # scaling '256x256_512x512'
        video = np.ascontiguousarray(video[:, :, :, ::-1])                                   # RGB -> BGR (run_MFT_tapvid.py:126)
        pts = d['points'] * np.array([W, H])
        return torch.from_numpy(video).cuda(), {'first': TV.sample_queries_first(d['occluded'], pts),
                                                'strided': TV.sample_queries_strided(d['occluded'], pts)}

    def run_all(seqs):
        n = 0
        for name in seqs:
            video, queries = prepare(name)
            cache = DeviceFlowCache(max_bytes=24 << 30)
            for mode in ('first', 'strided'):
                tracks, occl, k = TV.run_sequence(tracker, video, queries[mode], mode, flow_cache=cache)
                assert np.isfinite(tracks).all()
                n += k
            run_all.hits += cache.hits
            run_all.misses += cache.misses
        return n
    run_all.hits = run_all.misses = 0
    run_all(mine[:1])                                       # warm-up: one full sequence (allocations, pinned pools, clocks)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.resume()
    run_all.hits = run_all.misses = 0
    launches0 = tracker.engine.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    n_frames_done = run_all(mine)
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tracker.engine.check_device()
    dev_ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([dev_ms, wall, float(n_frames_done), float(run_all.hits), float(run_all.misses)], dtype=torch.float64, device='cuda')
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dev_ms, wall = mx[0].item(), mx[1].item()
    total_frames, hits, misses = tt[2].item(), tt[3].item(), tt[4].item()
    if rank == 0:
        line = {
            'metric': 'dense-track frames/sec', 'value': total_frames / (dev_ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world,
            'steps': int(total_frames), 'warmup': 1, 'ms_per_step': dev_ms / max(1.0, total_frames) * world, 'higher_is_better': True,
            'scaling': 'weak' if args.sequences <= world else 'strong', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': f'TAP-Vid style evaluation: {len(names)} synthetic sequences x {n_frames} frames (TAP-Vid pickle schema, 256x256 -> {W}x{H}), '
                                   'query modes first + strided (forward + backward from every 5th frame), deltas [inf,1,2,4,8,16,32], 12 GRU iters, '
                                   'device flow cache, point queries sampled on the device',
                       'sharding': 'sequences round-robin over ranks, no data-path collective', 'weights': wsrc,
                       'unit_of_work': 'a frame passed to MFT.init / MFT.track (cached pairs are not recomputed, like the reference)'},
            'e2e': {'value': total_frames / wall, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 32 * 3 * 4,
                    'note': 'wall clock of the whole evaluation loop incl. host bookkeeping; videos uploaded once per sequence'},
            'gpu_launches': int(tracker.engine.launch_count() - launches0),
            'flow_cache': {'hits': int(hits), 'misses': int(misses)},
            'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_demo(args):
    """BASELINE config 2: demo.py's loop over the whole demo video at 512x512 (init on frame 0, track every other frame,
    pageable numpy frames in, CPU results out, point queries converted per frame like demo.py:59-69)."""
    import torch
    from mft_b200.point_tracking import convert_to_point_tracking
    from mft_b200.synth import demo_video_frames
    H, W = parse_size(args.size, 512)
    frames = demo_video_frames((W, H))
    if len(frames) < 3:
        print(json.dumps({'metric': 'dense-track frames/sec', 'unavailable': 'demo video not reachable on this box'}), flush=True)
        return
    torch.cuda.set_device(0)
    weights, wsrc = load_weights()
    tracker = make_tracker(weights)
    yy, xx = np.mgrid[20:H - 10:30, 20:W - 10:30]
    queries = torch.from_numpy(np.stack([xx.ravel(), yy.ravel()], 1).astype(np.float32)).cuda()

    def run(device_frames):
        src = [torch.from_numpy(f).cuda() for f in frames] if device_frames else frames
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        tracker.init(src[0])
        for f in src[1:]:
            meta = tracker.track(f, device_result=device_frames)
            if not device_frames:
                convert_to_point_tracking(meta.result, queries)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3, time.perf_counter() - t0, meta

    run(True)                                                    # warm-up pass (allocations, clocks)
    sampler = ClockSampler(0)
    sampler.start()
    sampler.resume()
    dev_s, _, _ = run(True)
    _, e2e_s, meta = run(False)
    clocks = sampler.stop()
    tracker.engine.check_device()
    n = len(frames)
    assert tuple(meta.result.flow.shape) == (2, H, W) and not meta.result.flow.is_cuda
    line = {
        'metric': 'dense-track frames/sec', 'value': n / dev_s, 'unit': 'frames/s', 'n_gpus': 1, 'steps': n, 'warmup': n,
        'ms_per_step': dev_s / n * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'demo video',
        'config': {'workload': f'demo_in video, all {n} frames resized to {W}x{H}, deltas [inf,1,2,4,8,16,32], 12 GRU iters: init + track of every frame '
                               '(the first 32 frames run fewer than 7 chains), as demo.py', 'weights': wsrc,
                   'cache': 'per-frame working set (7 correlation pyramids, 300 MB) exceeds the L2; no flush',
                   'e2e': 'pageable numpy frames in, CPU FlowOUTrackingResult out, 272 point queries converted per frame (demo.py:59-69)'},
        'e2e': {'value': n / e2e_s, 'unit': 'frames/s', 'h2d_bytes_per_step': H * W * 3, 'd2h_bytes_per_step': 16 * H * W},
        'gpu_launches': int(tracker.engine.launch_count()), 'clocks': clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='track', choices=['track', 'flow-shard', 'tapvid', 'demo'])
    ap.add_argument('--size', default='', help='frame size: N or HxW (default 512; 1024 for --mode flow-shard)')
    ap.add_argument('--iters', type=int, default=0, help='GRU iterations for --mode flow-shard (default 32)')
    ap.add_argument('--sequences', type=int, default=8, help='--mode tapvid: number of synthetic sequences')
    ap.add_argument('--frames', type=int, default=24, help='--mode tapvid: frames per sequence')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-identity-check', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.mode == 'flow-shard':
        run_flow_shard(args)
    elif args.mode == 'tapvid':
        run_tapvid(args)
    elif args.mode == 'demo':
        run_demo(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
