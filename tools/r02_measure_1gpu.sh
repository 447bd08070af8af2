set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --steps 20 --warmup 5 --min-seconds 2.5 --no-ref-batch --no-cpu-baseline > $O/bench_cfg2_sustained.json 2> $O/err2
timeout 300 python bench.py --config cfg3 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/err3
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/err4
timeout 900 bash tools/cfg5_sweep.sh > $O/cfg5_sweep_1gpu.jsonl 2> $O/err5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-ref-batch --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psf_gen_bwd_tc2 -s 3 -c 1 -f -o $O/gen_bwd_tc2 python bench.py --steps 2 --warmup 1 --no-ref-batch --no-cpu-baseline > $O/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psf_gen_fwd_tc -s 3 -c 1 -f -o $O/gen_fwd_tc python bench.py --steps 2 --warmup 1 --no-ref-batch --no-cpu-baseline > $O/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psf_chain2_bwd -s 3 -c 1 -f -o $O/chain2_bwd python bench.py --steps 2 --warmup 1 --no-ref-batch --no-cpu-baseline > $O/ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psf_chain2_fwd -s 3 -c 1 -f -o $O/chain2_fwd python bench.py --steps 2 --warmup 1 --no-ref-batch --no-cpu-baseline > $O/ncu4.log 2>&1
ls -la $O
