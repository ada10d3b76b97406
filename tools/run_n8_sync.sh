#!/bin/bash
# 8-GPU A/B: FlatGradSync (one all-reduce after backward) vs torch DDP buckets; then the other workloads with the winner
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
run() {
  local name=$1; shift
  $TR bench.py --gpus 8 --steps 12 --warmup 4 "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT/$name.json") if l.startswith('{"metric"')][-1]
    print("$name", round(d["value"], 1), "sections/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], "loss", d["loss"], "copied", d["config"].get("grad_sync_copied_params"))
except Exception as e:
    print("$name FAILED", e)
PY
}
run r02g_n8_cfg2_flat --grad-sync flat
run r02g_n8_cfg2_ddp --grad-sync ddp
run r02g_n8_cfg2_flat2 --grad-sync flat
$TR bench.py --gpus 8 --grad-sync flat --timeline $OUT/r02g_timeline_n8_flat.txt > /dev/null 2> $OUT/r02g_timeline_n8.err
head -3 $OUT/r02g_timeline_n8_flat.txt.rank3
