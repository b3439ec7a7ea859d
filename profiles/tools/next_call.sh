#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 > $O/r2_bench_c3_mode1.json 2> $O/r2_bench_c3_mode1.err
for f in c3_mode0 c3_mode1; do echo "== $f"; grep -E "trace|c3:|Error" $O/r2_bench_$f.err | tail -3; head -c 300 $O/r2_bench_$f.json; echo; done
