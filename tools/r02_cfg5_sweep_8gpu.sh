N=8; O=gpurun_out/r02f_n8; mkdir -p $O
for C in 8 32 128; do for res in 1 0; do for lg in 9 12 14; do
    n=$((1 << lg)); L=$((lg + 1)); B=$(( (256 * 1024 * 1024) / (4 * lg * n * L) )); B=$(( B < 33 ? 33 : B ))
    cap=$(( (2 * 1024 * 1024 * 1024) / (4 * n * C) )); B=$(( B > cap ? cap : B ))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29614 + lg + C)) tools/chain_bench.py --seq-len $n --B $B --C $C --residual $res --iters 5 2>/dev/null | grep '^{'
done; done; done > $O/cfg5_sweep.jsonl
wc -l $O/cfg5_sweep.jsonl
