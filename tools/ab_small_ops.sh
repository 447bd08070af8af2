run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'])"; }
PSF_HEAD_KERNELS=0 PSF_INIT_LINEAR_KERNELS=0 run none
PSF_HEAD_KERNELS=0 run embed_only
PSF_INIT_LINEAR_KERNELS=0 run head_only
run both
PSF_HEAD_KERNELS=0 PSF_INIT_LINEAR_KERNELS=0 run none_again
