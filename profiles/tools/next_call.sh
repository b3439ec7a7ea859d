#!/bin/bash
# 8-GPU box: world 8 only (sharded check, sharded C4 bench, replicated C2 bench, C5 frames).
O=gpurun_out
bash profiles/tools/run_multi.sh 8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29777 bench.py --workload c5 --gpus 8 --steps 60 --warmup 3 --no-cpu > $O/r2_bench_c5_w8.json 2> $O/r2_bench_c5_w8.err
head -c 300 $O/r2_bench_replicas_w8.json; echo; head -c 300 $O/r2_bench_c5_w8.json; echo; grep -E "Error" $O/r2_bench_c5_w8.err | tail -2
