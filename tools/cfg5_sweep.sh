#!/bin/bash
# BASELINE cfg5: chord-chain sweep over n = 2^7 .. 2^16 x channels C in {8, 32, 128} x residual on / off.
# One JSON line per shape (tools/chain_bench.py): kernel generation, fwd and bwd ms, algorithmic GB/s and the fraction of the
# measured HBM copy peak.  Batch sized for >= 256 MB of W (and at least 33 = one wave of resident clusters), capped so that
# the [B, n, C] state stays under 2 GB.
#   tools/cfg5_sweep.sh            one GPU
#   under torchrun every rank runs the same sweep on its own GPU (replicas: the chain has no cross-GPU step); rank 0 prints
for C in 8 32 128; do
  for res in 1 0; do
    for lg in 7 9 11 12 13 14 16; do
      n=$((1 << lg)); L=$((lg + 1))
      B=$(( (256 * 1024 * 1024) / (4 * lg * n * L) )); B=$(( B < 33 ? 33 : B )); B=$(( B > 4096 ? 4096 : B ))
      cap=$(( (2 * 1024 * 1024 * 1024) / (4 * n * C) )); B=$(( B > cap ? cap : B )); B=$(( B < 1 ? 1 : B ))
      timeout 200 python tools/chain_bench.py --n $n --B $B --C $C --residual $res --iters 5 2>&1 | tail -1
    done
  done
done
