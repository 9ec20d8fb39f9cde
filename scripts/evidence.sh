#!/bin/bash
# Round evidence on one B200 (run under gpurun): bench line (the driver's command), the CPU reference arm, the
# reference's own CUDA, ncu launch list and one full capture of the two hot kernels.  Every ncu command follows the
# identical command exiting 0 without ncu; numbers printed under ncu are never bench values.
# usage: scripts/evidence.sh <prefix>   -> gpurun_out/<prefix>_*
set -u
P=${1:-r2x}
O=gpurun_out
mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${P}_bench_n1_config2.json 2> $O/${P}_bench.err || { echo "bench failed"; tail -5 $O/${P}_bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_reference_arm.json 2>> $O/${P}_bench.err
python bench.py --impl reference-gpu --steps 2 --warmup 1 > $O/${P}_bench_reference_gpu_arm.json 2>> $O/${P}_bench.err
python bench.py --particles 4096 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"polar_|ccf_|finalize|rotsum|mask_normalize|class_fsc|filter_center" --csv \
    --log-file $O/${P}_launches.csv python bench.py --particles 4096 --steps 1 --warmup 3 --no-cpu-baseline > $O/${P}_ncu_l.log 2>&1
python scripts/gpu_probe.py 2048 50 > $O/${P}_probe.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"polar_group|ccf_tm" -c 2 -o $O/${P}_full -f \
    python scripts/gpu_probe.py 2048 50 > $O/${P}_ncu_f.log 2>&1
python - <<PY
import json
d=json.loads(open("$O/${P}_bench_n1_config2.json").read().strip().splitlines()[-1])
print("bench: value %.4e e2e %.4e ms/step %.1f full-iteration %.3f s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["s_per_full_iteration"]), d["stage_ms_per_step"], "roofline frac %.4f" % d["roofline"]["frac"])
for f in ("reference_arm", "reference_gpu_arm"):
    r=json.loads(open("$O/${P}_bench_%s.json" % f).read().strip().splitlines()[-1]); print(f, r.get("value"), r.get("unavailable"))
PY
