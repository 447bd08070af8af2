#!/bin/bash
# BASELINE cfg5: chord-chain micro-sweep n = 2^7 .. 2^16 (C = 8, residual on), batch sized for >= 256 MB of W.
# One JSON line per shape (tools/chain_bench.py): path / kernel generation, fwd and bwd ms, algorithmic GB/s.
for lg in 7 8 9 10 11 12 13 14 15 16; do
  n=$((1 << lg)); L=$((lg + 1))
  B=$(( (256 * 1024 * 1024) / (4 * lg * n * L) )); B=$(( B < 33 ? 33 : B )); B=$(( B > 4096 ? 4096 : B ))
  timeout 200 python tools/chain_bench.py --n $n --B $B --iters 5 2>&1 | tail -1
done
