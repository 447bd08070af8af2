#!/bin/bash
# A/B of differently tuned builds (sparsefactorization_b200/_lib/variants/*.so): generator-only bench of each.
echo "default :: $(timeout 200 python tools/gen_bench.py "$@" 2>&1 | tail -1)"
for so in sparsefactorization_b200/_lib/variants/*.so; do
  echo "$(basename $so) :: $(PSF_B200_LIB=$so timeout 200 python tools/gen_bench.py "$@" 2>&1 | tail -1)"
done
