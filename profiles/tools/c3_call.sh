#!/bin/bash
# Streaming mapper after the reservation change: tests, allocation cost on the box, C3 in both matching modes, the headline line (traffic stamped), K = 8 contexts.
set +e
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 120 python profiles/tools/alloc_cost.py > $O/r2_alloc_cost.txt 2>&1; cat $O/r2_alloc_cost.txt
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 > $O/r2_bench_c3_mode1.json 2> $O/r2_bench_c3_mode1.err
timeout 400 python bench.py --steps 200 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
timeout 200 python bench.py --steps 240 --warmup 5 --contexts 8 --no-cpu > $O/r2_bench_n1_k8.json 2>> $O/r2_bench_n1.err
tail -c 300 $O/r2_bench_n1.err; head -c 250 $O/r2_bench_n1.json; echo; head -c 250 $O/r2_bench_n1_k8.json; echo
for f in c3_mode0 c3_mode1; do echo "== $f"; grep -E "c3:|Error" $O/r2_bench_$f.err | tail -2; head -c 300 $O/r2_bench_$f.json; echo; done
