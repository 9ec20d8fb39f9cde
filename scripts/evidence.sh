#!/bin/bash
# Round evidence on one B200 (run under gpurun): bench line, reference arm, ncu launch list and one full capture
# of the two hot kernels.  Every ncu command follows the identical command exiting 0 without ncu; numbers printed
# under ncu are never bench values.  usage: scripts/evidence.sh <prefix>   -> gpurun_out/<prefix>_*
set -u
P=${1:-r1x}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err || { echo "bench failed"; tail -5 $O/${P}_bench.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $O/${P}_ref.json 2>> $O/${P}_bench.err
python bench.py --particles 4096 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"polar_|ccf_|finalize|rotsum|mask_normalize" --csv \
    --log-file $O/${P}_launches.csv python bench.py --particles 4096 --steps 1 --warmup 3 --no-cpu-baseline > $O/${P}_ncu_l.log 2>&1
python scripts/gpu_probe.py 2048 50 > $O/${P}_probe.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"polar_group|ccf_tm" -c 2 -o $O/${P}_full -f \
    python scripts/gpu_probe.py 2048 50 > $O/${P}_ncu_f.log 2>&1
tail -c 400 $O/${P}_bench.json; echo; grep ms_ccf $O/${P}_probe.log | tail -1
