#!/bin/bash
# One-GPU runs of the other configs: C3 (1000-scan stream, both matching modes), C5 at N=1, the C4 map on one GPU (strong-scaling baseline), the CPU arm.
set +e
O=gpurun_out
timeout 700 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
timeout 700 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 > $O/r2_bench_c3_mode1.json 2> $O/r2_bench_c3_mode1.err
timeout 400 python bench.py --workload c5 --steps 60 --warmup 3 > $O/r2_bench_c5_w1.json 2> $O/r2_bench_c5_w1.err
timeout 500 python bench.py --mode sharded --steps 100 --warmup 5 --cpu-scans 2 > $O/r2_bench_sharded_w1.json 2> $O/r2_bench_sharded_w1.err
timeout 400 python bench.py --impl reference --steps 200 --warmup 5 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
for f in c3_mode0 c3_mode1 c5_w1 sharded_w1 reference_arm; do echo "== $f"; tail -c 300 $O/r2_bench_$f.err; head -c 260 $O/r2_bench_$f.json; echo; done
