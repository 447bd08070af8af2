#!/bin/bash
# sweep of the fused-path launch knobs (environment overrides read by psf_fused_plan)
run() { echo "$1 :: $(env $1 timeout 90 python tools/chain_bench.py --B ${2:-264} --iters 5 2>&1 | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print(r['plan']['cluster'], r['plan']['rows_per_cta'], r['plan']['threads'], r['plan']['resident_clusters'], 'fwd', r['fwd_ms'], 'bwd', r['bwd_ms'], 'GB/s', r['fwdbwd_GBps'])")"; }
run "PSF_FUSED_STAGES=2 PSF_FUSED_STAGES_BWD=2"
run "PSF_FUSED_STAGES=3 PSF_FUSED_STAGES_BWD=2"
run "PSF_FUSED_STAGES=2 PSF_FUSED_STAGES_BWD=3"
run "PSF_FUSED_CONSUMER_WARPS=4"
run "PSF_FUSED_MAX_ROWS=4096"
