#!/bin/bash
# final-tree 8-GPU lines + device timeline of rank 0 / rank 7 (one node, torchrun, NCCL)
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 8 --steps 15 --warmup 4 > $OUT/r02e_n8_cfg2_b16.json 2> $OUT/r02e_n8_cfg2_b16.err
$TR bench.py --gpus 8 --timeline $OUT/r02e_timeline_n8.txt > /dev/null 2> $OUT/r02e_timeline_n8.err
python bench.py --gpus 1 --steps 15 --warmup 4 --no-cpu-baseline > $OUT/r02e_n1_samebox.json 2> /dev/null
python - <<PY
import json
for f in ("r02e_n8_cfg2_b16", "r02e_n1_samebox"):
    try:
        d = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "sections/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), d["clocks"])
    except Exception as e:
        print(f, "FAILED", e)
PY
head -4 $OUT/r02e_timeline_n8.txt
