#!/bin/bash
# final single-GPU bench lines of the round (after the small-operator kernels and the last generator changes)
O=gpurun_out/r02h; mkdir -p $O
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/err2
timeout 300 python bench.py --steps 20 --warmup 5 --min-seconds 2.5 --no-ref-batch --no-cpu-baseline > $O/bench_cfg2_sustained.json 2>> $O/err2
timeout 300 python bench.py --config cfg3 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg3.json 2>> $O/err2
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg4.json 2>> $O/err2
timeout 120 python bench.py > $O/bench_default.json 2>> $O/err2
tail -c 400 $O/err2
