#!/bin/bash
O=gpurun_out
timeout 500 python bench.py --impl reference --contexts 16 --steps 128 --warmup 16 > $O/r2_bench_reference_arm_k16.json 2> $O/r2_bench_reference_arm_k16.err
tail -c 200 $O/r2_bench_reference_arm_k16.err; head -c 300 $O/r2_bench_reference_arm_k16.json; echo
