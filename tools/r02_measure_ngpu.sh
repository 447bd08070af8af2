#!/bin/bash
# N-GPU measurement batch (weak scaling: one process per GPU, fixed batch per GPU).  usage: tools/r02_measure_ngpu.sh N
N=${1:-2}
O=gpurun_out/r02f_n$N
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/err_cfg2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --config cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-batch > $O/bench_cfg3.json 2> $O/err_cfg3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-batch > $O/bench_cfg4.json 2> $O/err_cfg4
if [ "$N" = "8" ]; then
  # cfg5 at 8 GPUs: the chain has no cross-GPU step, every rank runs the same sweep on its own GPU at the same time
  for C in 8 32 128; do for res in 1 0; do for lg in 9 12 14; do
    n=$((1 << lg)); L=$((lg + 1)); B=$(( (256 * 1024 * 1024) / (4 * lg * n * L) )); B=$(( B < 33 ? 33 : B ))
    cap=$(( (2 * 1024 * 1024 * 1024) / (4 * n * C) )); B=$(( B > cap ? cap : B ))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/chain_bench.py --seq-len $n --B $B --C $C --residual $res --iters 5 2>/dev/null | grep '^{'
  done; done; done > $O/cfg5_sweep.jsonl
fi
tail -c 600 $O/bench_cfg2.json; tail -c 300 $O/err_cfg2
