#!/bin/bash
# Round-2 evidence pass on one B200 (run under gpurun): bench, reference arm, ncu launch list, ncu --set full of
# the streaming kernel (cfg4), the resident kernel (cfg1), the batched shape (cfg5) and the L2-pinned path (cfg2),
# latency table, per-CTA straggler dump.  Every ncu capture runs only after the same command exited 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
P=r2
set -x
python bench.py --steps 20 --warmup 3 > $O/${P}_bench.json 2> $O/${P}_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $O/${P}_bench_reference.json 2> $O/${P}_bench_reference.err
python scripts/latency.py > $O/${P}_latency.txt 2>&1
NLO_DEBUG_TIMES=1 python scripts/phase_times.py >> $O/${P}_latency.txt 2>&1
python scripts/phase_dump.py 8000000 ndt6 $O/${P}_dump_8M.bin > /dev/null 2>&1
python scripts/phase_dump.py 100000 ndt6 $O/${P}_dump_100k.bin > /dev/null 2>&1
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra --no-dropin"
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_64M.csv $B > /dev/null 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gn_iteration_kernel -s 1 -c 1 -f -o $O/${P}_full_64M $B > $O/${P}_ncu_full.log 2>&1
ncu -i $O/${P}_full_64M.ncu-rep --page raw --csv > $O/${P}_ncu_full_64M_raw.csv 2>/dev/null
C1="python scripts/one_solve.py 100000 ndt6 40"
$C1 && ncu --set full --clock-control none -k regex:gn_resident_kernel -s 1 -c 1 -f -o $O/${P}_full_cfg1 $C1 > /dev/null 2>&1
ncu -i $O/${P}_full_cfg1.ncu-rep --page raw --csv > $O/${P}_ncu_full_cfg1_resident_raw.csv 2>/dev/null
C2="python scripts/one_solve.py 1000000 ndt3 40"
$C2 && ncu --set full --clock-control none -k regex:gn_iteration_kernel -s 1 -c 1 -f -o $O/${P}_full_cfg2 $C2 > /dev/null 2>&1
ncu -i $O/${P}_full_cfg2.ncu-rep --page raw --csv > $O/${P}_ncu_full_cfg2_l2_raw.csv 2>/dev/null
C5="python scripts/batched.py 4096 20000"
$C5 > $O/${P}_cfg5.txt 2>&1 && ncu --set full --clock-control none -k regex:gn_iteration_kernel -s 1 -c 1 -f -o $O/${P}_full_cfg5 $C5 > /dev/null 2>&1
ncu -i $O/${P}_full_cfg5.ncu-rep --page raw --csv > $O/${P}_ncu_full_cfg5_batched_raw.csv 2>/dev/null
rm -f $O/${P}_full_cfg1.ncu-rep $O/${P}_full_cfg2.ncu-rep $O/${P}_full_cfg5.ncu-rep
python scripts/registration_fixture.py > $O/${P}_registration_fixture.json 2>&1
python scripts/map_hash.py > $O/${P}_map_hash.json 2>&1
ls -la $O | grep ${P}_
