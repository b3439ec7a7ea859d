#!/bin/bash
# Final one-GPU evidence of the round on the final sources: tests, headline line, throughput mode, solver phases, per-kernel table, ncu --set full of the two hot
# kernels (their summaries are stamped with the source sha1 so that bench.py's roofline.traffic matches the shipped kernels), C3 streams in both matching modes.
set +e
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 400 python bench.py --steps 200 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
timeout 200 python bench.py --steps 240 --warmup 5 --contexts 2 --no-cpu > $O/r2_bench_n1_k2.json 2>> $O/r2_bench_n1.err
timeout 200 python bench.py --steps 240 --warmup 5 --contexts 4 --no-cpu > $O/r2_bench_n1_k4.json 2>> $O/r2_bench_n1.err
timeout 100 python profiles/tools/solver_phases.py > $O/r2_solver_phases.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -c 900 --csv --log-file $O/r2_kernels.csv python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2_ncu_kernels.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_blocks_kernel -s 12 -c 4 -f -o $O/ncu_knn_blocks_r2 python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2_ncu_knn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lm_solve_kernel -s 12 -c 4 -f -o $O/ncu_lm_solve_r2 python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2_ncu_solve.log 2>&1
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 > $O/r2_bench_c3_mode1.json 2> $O/r2_bench_c3_mode1.err
tail -c 300 $O/r2_bench_n1.err; head -c 250 $O/r2_bench_n1.json; echo; tail -2 $O/r2_solver_phases.txt | cut -c1-400
for f in c3_mode0 c3_mode1; do echo "== $f"; grep -E "c3:|Error" $O/r2_bench_$f.err | tail -2; head -c 400 $O/r2_bench_$f.json; echo; done
