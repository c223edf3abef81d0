#!/bin/bash
# Multi-GPU evidence (run on an N-GPU box):   gpurun --gpus 2 --timeout 1200 -- 'bash tools/multi_gpu_evidence.sh r2e 2'
# Writes under gpurun_out/: the GPU test log, the bit-identity checks of the two sharded trackers (SURVEY 8e ii/iii), and
# the bench lines of BASELINE configs 3-5 (flow-shard at 1 and N GPUs, TAP-Vid style evaluation, 1080p).
tag=${1:-r2e}; n=${2:-2}
mkdir -p gpurun_out
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_pytest.txt
$tr tools/flow_shard_check.py > gpurun_out/${tag}_flow_shard_check_${n}gpu.log 2>&1; tail -2 gpurun_out/${tag}_flow_shard_check_${n}gpu.log
$tr tools/delta_shard_check.py > gpurun_out/${tag}_delta_shard_check_${n}gpu.log 2>&1; tail -2 gpurun_out/${tag}_delta_shard_check_${n}gpu.log
CHECK_SIZE=512 $tr tools/delta_shard_check.py > gpurun_out/${tag}_delta_shard_check_512_${n}gpu.log 2>&1; tail -2 gpurun_out/${tag}_delta_shard_check_512_${n}gpu.log
python bench.py --mode flow-shard --gpus 1 --steps 8 > gpurun_out/${tag}_flowshard_1gpu.json 2> gpurun_out/${tag}_flowshard_1gpu.err; tail -c 600 gpurun_out/${tag}_flowshard_1gpu.json
$tr bench.py --mode flow-shard --gpus $n --steps 8 > gpurun_out/${tag}_flowshard_${n}gpu.json 2> gpurun_out/${tag}_flowshard_${n}gpu.err; tail -c 600 gpurun_out/${tag}_flowshard_${n}gpu.json
$tr bench.py --mode tapvid --gpus $n --sequences 4 --frames 24 > gpurun_out/${tag}_tapvid_${n}gpu.json 2> gpurun_out/${tag}_tapvid_${n}gpu.err; tail -c 600 gpurun_out/${tag}_tapvid_${n}gpu.json
python bench.py --size 1080x1920 --steps 10 --no-cpu-baseline > gpurun_out/${tag}_1080p_1gpu.json 2> gpurun_out/${tag}_1080p_1gpu.err; tail -c 600 gpurun_out/${tag}_1080p_1gpu.json
