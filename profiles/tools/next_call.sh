#!/bin/bash
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
timeout 500 python bench.py --impl reference --contexts 3 --steps 60 --warmup 3 > $O/r2_bench_reference_arm_k3.json 2> $O/r2_bench_reference_arm_k3.err
tail -c 200 $O/r2_bench_reference_arm_k3.err; head -c 400 $O/r2_bench_reference_arm_k3.json; echo
