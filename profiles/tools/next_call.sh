#!/bin/bash
# The command of the next gpurun call (edited between calls; the snapshot is taken when the call gets its box).
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "init_gate or shard_build or golden or frame_to_pose or scene" 2>&1 | tail -15 > $O/r2_pytest_new.log; tail -4 $O/r2_pytest_new.log
bash profiles/tools/run_multi.sh 2 "--cpu-scans 2" skip
bash profiles/tools/run_multi.sh 4 "" skip
timeout 500 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
echo "== c3_mode0"; grep -E "c3:|Error" $O/r2_bench_c3_mode0.err | tail -3; head -c 400 $O/r2_bench_c3_mode0.json; echo
