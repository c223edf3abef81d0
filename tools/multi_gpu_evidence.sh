#!/bin/bash
# Multi-GPU evidence (run on an N-GPU box):   gpurun --gpus 8 --timeout 1200 -- 'bash tools/multi_gpu_evidence.sh r2k 8'
# Writes under gpurun_out/: (N = 2 only: the GPU test log,) the bit-identity checks of the two sharded trackers (SURVEY 8e
# ii/iii) and the bench lines of BASELINE configs 3-5 at N GPUs: per-timestep flow sharding of ONE 1024x1024 / 32-iteration
# video (strong scaling; the 1-GPU line is written too when N = 2), TAP-Vid style evaluation with sequences round-robin over the
# ranks, one 1080p sequence per GPU.
tag=${1:-r2e}; n=${2:-2}
mkdir -p gpurun_out
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
if [ "$n" = "2" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_pytest.txt
  python bench.py --mode flow-shard --gpus 1 --steps 8 > gpurun_out/${tag}_flowshard_1gpu.json 2> gpurun_out/${tag}_flowshard_1gpu.err; tail -c 300 gpurun_out/${tag}_flowshard_1gpu.json
  python bench.py --size 1080x1920 --steps 10 --no-cpu-baseline > gpurun_out/${tag}_1080p_1gpu.json 2> gpurun_out/${tag}_1080p_1gpu.err; tail -c 300 gpurun_out/${tag}_1080p_1gpu.json
fi
$tr tools/flow_shard_check.py > gpurun_out/${tag}_flow_shard_check_${n}gpu.log 2>&1; tail -1 gpurun_out/${tag}_flow_shard_check_${n}gpu.log
CHECK_SIZE=512 $tr tools/delta_shard_check.py > gpurun_out/${tag}_delta_shard_check_512_${n}gpu.log 2>&1; tail -1 gpurun_out/${tag}_delta_shard_check_512_${n}gpu.log
$tr bench.py --mode flow-shard --gpus $n --steps 16 > gpurun_out/${tag}_flowshard_${n}gpu.json 2> gpurun_out/${tag}_flowshard_${n}gpu.err; tail -c 300 gpurun_out/${tag}_flowshard_${n}gpu.json; grep identity gpurun_out/${tag}_flowshard_${n}gpu.err
$tr bench.py --mode tapvid --gpus $n --sequences $((n > 4 ? n : 4)) --frames 24 > gpurun_out/${tag}_tapvid_${n}gpu.json 2> gpurun_out/${tag}_tapvid_${n}gpu.err; tail -c 300 gpurun_out/${tag}_tapvid_${n}gpu.json
$tr bench.py --size 1080x1920 --gpus $n --steps 10 > gpurun_out/${tag}_1080p_${n}gpu.json 2> gpurun_out/${tag}_1080p_${n}gpu.err; tail -c 300 gpurun_out/${tag}_1080p_${n}gpu.json
