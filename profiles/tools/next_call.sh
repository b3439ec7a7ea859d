#!/bin/bash
# final evidence on the final sources: tests, ncu --set full of the kNN kernel (stamp = sha1 of knn.cu), headline line with both traffic figures, C3 in both modes
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_blocks_kernel -s 12 -c 4 -f -o $O/ncu_knn_blocks_r2 python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2_ncu_knn.log 2>&1
python profiles/summarize.py $O/ncu_knn_blocks_r2.ncu-rep profiles/ncu_knn_blocks_kernel_r2.txt "knn_blocks_kernel (round 2: depth-first nearest-child-first, lane-distributed top-5), 4 consecutive launches of the bench" knn.cu > /dev/null 2>&1
timeout 400 python bench.py --steps 200 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 0 > $O/r2_bench_c3_mode0.json 2> $O/r2_bench_c3_mode0.err
timeout 600 python bench.py --workload c3 --steps 995 --warmup 5 --matching-mode 1 > $O/r2_bench_c3_mode1.json 2> $O/r2_bench_c3_mode1.err
tail -c 300 $O/r2_bench_n1.err; head -c 250 $O/r2_bench_n1.json; echo
for f in c3_mode0 c3_mode1; do echo "== $f"; grep -E "trace|c3:|Error" $O/r2_bench_$f.err | tail -3; head -c 300 $O/r2_bench_$f.json; echo; done
