#!/usr/bin/env bash
# 2-GPU probe: does leaving SMs to NCCL (CRVQA_GEMM_SMS caps the persistent GEMM grids) and / or limiting NCCL's CTAs
# (NCCL_MAX_CTAS) help the data-parallel step?  Measured on 2 x B200 (round 2): base 14.20 ms/step; GEMM_SMS=140 14.58;
# 140 + MAX_CTAS=8 14.28; 144 + MAX_CTAS=4 15.13; MAX_CTAS=4 alone 16.61; 132 + MAX_CTAS=16 14.17 -- no setting beats
# the default (all SMs to the GEMMs, NCCL's own CTA count), so neither is set by the engine.
#   gpurun --gpus 2 -- bash tests/sms_sweep.sh
run() { # label, env...
  label="$1"; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm', round(d['roofline']['achieved']))"
}
run base X=1
run sms140 CRVQA_GEMM_SMS=140
run sms140_cta8 CRVQA_GEMM_SMS=140 NCCL_MAX_CTAS=8
run sms144_cta4 CRVQA_GEMM_SMS=144 NCCL_MAX_CTAS=4
run cta4 NCCL_MAX_CTAS=4
run sms132_cta16 CRVQA_GEMM_SMS=132 NCCL_MAX_CTAS=16
