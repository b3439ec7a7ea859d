#!/bin/bash
# after the one-line racecheck fix in solve.cu: tests, racecheck + memcheck over smoke, ncu --set full of the solver (stamp), headline line
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout -k 5 120 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r2_sanitizer_racecheck.log 2>&1; tail -2 $O/r2_sanitizer_racecheck.log
timeout -k 5 120 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r2_sanitizer_memcheck.log 2>&1; tail -1 $O/r2_sanitizer_memcheck.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lm_solve_kernel -s 12 -c 4 -f -o $O/ncu_lm_solve_r2 python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2_ncu_solve.log 2>&1
python profiles/summarize.py $O/ncu_lm_solve_r2.ncu-rep profiles/ncu_lm_solve_kernel_r2.txt "lm_solve_kernel (round 2: master-less exchange, speculative LM step, distributed K10), 4 consecutive launches of the bench" solve.cu > /dev/null 2>&1
timeout 300 python bench.py --steps 200 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
tail -c 200 $O/r2_bench_n1.err; head -c 250 $O/r2_bench_n1.json; echo
