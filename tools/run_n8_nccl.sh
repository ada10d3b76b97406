#!/bin/bash
# 8-GPU A/B of NCCL settings for the gradient all-reduce (one node, torchrun).  Which algorithm NCCL picks by itself is in
# the NCCL_DEBUG=INFO log of the first run.
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
run() {
  local name=$1; shift; local envs=$1; shift
  env $envs $TR bench.py --gpus 8 --steps 12 --warmup 4 "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"], 1), "sections/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run r02f_n8_default "NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING"
grep -E "NVLS|Algo|algorithm|Channel.*NVLS|nvls" $OUT/r02f_n8_default.err | head -12 | cut -c1-200
run r02f_n8_nvls "NCCL_ALGO=NVLS"
run r02f_n8_tree "NCCL_ALGO=Tree"
run r02f_n8_ring_simple "NCCL_ALGO=Ring NCCL_PROTO=Simple"
