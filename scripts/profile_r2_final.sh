#!/bin/bash
# Re-capture of the ncu evidence on the last code of round 2 (the streaming kernel with two warp groups per CTA):
# launch list and `--set full` of the 64 M-point persistent launch, `--set full` of cfg2.  Run under gpurun; every
# capture only after the same command exited 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
P=r2w
set -x
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra --no-dropin"
$B > $O/${P}_bench_5steps.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_64M.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_iteration_kernel -s 1 -c 1 -f -o $O/${P}_full_64M $B > $O/${P}_ncu_full.log 2>&1
ncu -i $O/${P}_full_64M.ncu-rep --page raw --csv > $O/${P}_ncu_full_64M_raw.csv 2>/dev/null
C2="python scripts/one_solve.py 1000000 ndt3 40"
$C2 && ncu --set full --clock-control none -k regex:gn_iteration_kernel -s 1 -c 1 -f -o $O/${P}_full_cfg2 $C2 > /dev/null 2>&1
ncu -i $O/${P}_full_cfg2.ncu-rep --page raw --csv > $O/${P}_ncu_full_cfg2_l2_raw.csv 2>/dev/null
rm -f $O/${P}_full_cfg2.ncu-rep $O/${P}_full_64M.ncu-rep
ls -la $O | grep ${P}_
