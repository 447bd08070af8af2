#!/bin/bash
# A/B of differently tuned builds of libpsf_b200.so (sparsefactorization_b200/_lib/variants/*.so, built with
# `python -m sparsefactorization_b200.build --out ... -DPSF2_...`): chain-only bench of each.
B=${1:-132}
line() { python -c "import json,sys; r=json.loads(sys.stdin.read()); print(r['plan']['resident_clusters'], 'fwd', r['fwd_ms'], 'bwd', r['bwd_ms'], 'GB/s', r['fwd_GBps'], r['bwd_GBps'], r['fwdbwd_GBps'], 'frac', r['frac_of_measured_peak'])"; }
echo "default :: $(timeout 120 python tools/chain_bench.py --B $B --iters 10 2>&1 | tail -1 | line)"
for so in sparsefactorization_b200/_lib/variants/*.so; do
  echo "$(basename $so) :: $(PSF_B200_LIB=$so timeout 120 python tools/chain_bench.py --B $B --iters 10 2>&1 | tail -1 | line)"
done
