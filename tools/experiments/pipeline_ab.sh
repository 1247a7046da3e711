#!/bin/bash
# gpurun --gpus 2 -- bash tools/experiments/pipeline_ab.sh : A/B of the pipelined optimizer pass (N = 2)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out; mkdir -p $O
T0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--gpus 2 --steps 10 --warmup 3 --no-extra --no-hbm-kernels --no-cpu-baseline --infer-iters 1 --infer-batch 2048"
timeout 120 $TR --master-port 29701 tools/experiments/chunked_nccl_check.py > $O/r2p_chunk_check.txt 2>&1; echo "chunk check rc=$? t=$((SECONDS-T0))"; grep "equal\|all-reduce of" $O/r2p_chunk_check.txt
timeout 200 $TR --master-port 29702 bench.py $B --pipeline-adam 1 > $O/r2p_bench_n2_pipe1.json 2> $O/r2p_bench_n2_pipe1.err; echo "pipe1 rc=$? t=$((SECONDS-T0))"
timeout 150 $TR --master-port 29703 bench.py $B --pipeline-adam 0 --no-dp-parity > $O/r2p_bench_n2_pipe0.json 2> $O/r2p_bench_n2_pipe0.err; echo "pipe0 rc=$? t=$((SECONDS-T0))"
if [ $((SECONDS-T0)) -lt 150 ]; then
ES_DP_PIPELINE_CHUNKS=8 timeout 120 $TR --master-port 29704 bench.py $B --pipeline-adam 1 --no-dp-parity > $O/r2p_bench_n2_pipe1_c8.json 2> $O/r2p_bench_n2_pipe1_c8.err; echo "pipe1 c8 rc=$? t=$((SECONDS-T0))"
fi
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2p_bench_n2_pipe*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('comm'), (d.get('dp_parity') or {}).get('replicas_identical'))
    except Exception as e:
        print(f, 'unreadable', e)
P
