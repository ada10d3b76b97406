#!/bin/bash
# 8-GPU bench lines (one node, torchrun, NCCL).  usage: tools/run_n8.sh <outdir>
OUT=${1:-gpurun_out}
run() {  # name, env, args...
  local name=$1; shift; local envs=$1; shift
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 10 --warmup 4 "$@" > $OUT/$name.json 2> $OUT/$name.err
  tail -c 400 $OUT/$name.json | head -c 10 > /dev/null
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"], 1), "sections/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), d["clocks"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run n8_cfg2_b16 "A=1"
run n8_cfg2_b16_maxctas8 "NCCL_MAX_CTAS=8"
run n8_cfg2_b16_bucket48 "MMGL_DDP_BUCKET_MB=48"
run n8_cfg4_b16 "A=1" --workload cfg4
run n8_cfg3_b16 "A=1" --workload cfg3
run n8_cfg5_b8 "A=1" --workload cfg5 --batch 8
