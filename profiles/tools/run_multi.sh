#!/bin/bash
# Multi-GPU evidence run (under gpurun --gpus >= N): sharded parity check, sharded (C4, strong scaling) and replicated (C2, weak scaling) bench lines.
# usage: run_multi.sh N [extra bench args for the sharded line, default --no-cpu] [skip-replicas]
set +e
N=$1; O=gpurun_out; P=$((29500 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port $P tests/multi_gpu/sharded_check.py > $O/r2_sharded_check_w$N.log 2>&1
timeout 500 $TR --master-port $((P+10)) bench.py --gpus $N --mode sharded --steps 100 --warmup 5 ${2:---no-cpu} > $O/r2_bench_sharded_w$N.json 2> $O/r2_bench_sharded_w$N.err
if [ -z "$3" ]; then timeout 400 $TR --master-port $((P+20)) bench.py --gpus $N --steps 200 --warmup 5 --no-cpu > $O/r2_bench_replicas_w$N.json 2> $O/r2_bench_replicas_w$N.err; fi
grep -E "SHARDED|shard sizes|case rep=0" $O/r2_sharded_check_w$N.log | tail -8; grep -E "Error|assert" $O/r2_bench_sharded_w$N.err | tail -3; head -c 300 $O/r2_bench_sharded_w$N.json; echo
