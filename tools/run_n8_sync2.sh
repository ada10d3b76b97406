#!/bin/bash
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
run() {
  local name=$1; shift; local envs=$1; shift
  env $envs $TR bench.py --gpus 8 --steps 12 --warmup 4 "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT/$name.json") if l.startswith('{"metric"')][-1]
    print("$name", round(d["value"], 1), "sections/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], "loss", d["loss"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run r02g_n8_cfg2_flat_sum "MMGL_FLAT_SYNC_OP=sum"
run r02g_n8_cfg2_flat_bf16 "MMGL_FLAT_SYNC_BF16=1"
run r02g_n8_cfg4_flat "A=1" --workload cfg4
run r02g_n8_cfg3_flat "A=1" --workload cfg3
