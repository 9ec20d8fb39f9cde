#!/bin/bash
# 8-GPU lines of the named configurations (run under `gpurun --gpus 8`): config 2 weak scaling (the driver's own
# SCALE run), config 4 strong scaling in both variants SURVEY 8(d) recommends, config 5 (named for 8 GPUs).
# usage: scripts/scale_n8.sh <prefix> [N]
P=${1:-r2}; N=${2:-8}; O=gpurun_out; mkdir -p $O
run() { tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline "$@" > $O/${P}_bench_n${N}_$tag.json 2> $O/${P}_bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${P}_bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("$tag N=$N value %.4e e2e %.4e ms/step %.1f full %.3f s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["s_per_full_iteration"]), d["stage_ms_per_step"])
except Exception as exc:
    print("$tag failed:", exc); print(open("$O/${P}_bench_n${N}_$tag.err").read()[-1500:])
PY
}
run config2_weak --config 2
run config4_strong --config 4 --scaling strong
run config4_ou56_strong --config 4 --variant ou56 --scaling strong
run config5_weak --config 5
