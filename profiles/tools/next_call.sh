#!/bin/bash
# The command of the next gpurun call (edited between calls; the snapshot is taken when the call gets its box).
# 4-GPU box: world 2 and world 4 (sharded check, sharded C4 bench, replicated C2 bench), then the C3 streams on one GPU.
bash profiles/tools/run_multi.sh 2 "--cpu-scans 2"
bash profiles/tools/run_multi.sh 4
O=gpurun_out
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 > $O/r2_bench_c3_mode1.json 2> $O/r2_bench_c3_mode1.err
for f in c3_mode0 c3_mode1; do echo "== $f"; tail -c 300 $O/r2_bench_$f.err; head -c 300 $O/r2_bench_$f.json; echo; done
