#!/bin/bash
O=gpurun_out
LL_MAPPER_TRACE=1 timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 --no-cpu --dump-poses $O/r2_c3_track_mode0.npy > $O/r2_bench_c3_mode0_trace.json 2> $O/r2_bench_c3_mode0_trace.err
LL_MAPPER_TRACE=1 timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 --no-cpu > $O/r2_bench_c3_mode1_trace.json 2> $O/r2_bench_c3_mode1_trace.err
for m in 0 1; do echo "== mode $m"; grep -E "trace|Error|c3:" $O/r2_bench_c3_mode${m}_trace.err | tail -8; head -c 200 $O/r2_bench_c3_mode${m}_trace.json; echo; done
