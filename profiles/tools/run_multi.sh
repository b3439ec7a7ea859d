#!/bin/bash
# Multi-GPU evidence run (under gpurun --gpus N): sharded parity check, sharded (C4, strong scaling) and replicated (C2, weak scaling) bench lines.
# usage: run_multi.sh N
set +e
N=$1; O=gpurun_out; P=$((29500 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port $P tests/multi_gpu/sharded_check.py > $O/r2_sharded_check_w$N.log 2>&1
timeout 500 $TR --master-port $((P+10)) bench.py --gpus $N --mode sharded --steps 100 --warmup 5 ${2:---no-cpu} > $O/r2_bench_sharded_w$N.json 2> $O/r2_bench_sharded_w$N.err
timeout 400 $TR --master-port $((P+20)) bench.py --gpus $N --steps 200 --warmup 5 --no-cpu > $O/r2_bench_replicas_w$N.json 2> $O/r2_bench_replicas_w$N.err
tail -4 $O/r2_sharded_check_w$N.log; tail -c 300 $O/r2_bench_sharded_w$N.err; head -c 400 $O/r2_bench_sharded_w$N.json; echo; head -c 200 $O/r2_bench_replicas_w$N.json
