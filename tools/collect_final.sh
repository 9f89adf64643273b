#!/bin/bash
# copy the final pass (gpurun_out/r2final) into profiles/ as the round's committed evidence
set -e
cd /root/repo
R=gpurun_out/r2final
for n in cfg0 cfg1 cfg2 cfg3 cfg4 cfg1_reference cfg1_sorted cfg1_v06; do tail -n 1 $R/bench_$n.json > profiles/r2_bench_$n.json; done
python tools/ncu_summary.py launches $R/launches.csv > profiles/r2_launches.txt
python tools/ncu_table.py $R/prof_update.ncu-rep > profiles/r2_ncu_update.txt
python tools/ncu_table.py $R/prof_map.ncu-rep > profiles/r2_ncu_map.txt
python tools/ncu_summary.py full $R/prof_update.ncu-rep profiles/kernel_traffic.json > profiles/r2_ncu_update_full.txt 2>&1
python tools/ncu_summary.py full $R/prof_map.ncu-rep /tmp/map_traffic.json > profiles/r2_ncu_map_full.txt 2>&1
grep -v "^Multi" $R/timeline.log > profiles/r2_timeline.txt
python tools/sass_table.py > profiles/r2_sass_opcodes.txt
tail -n 2 $R/pytest_gpu.log > profiles/r2_pytest_gpu.txt
cat $R/smoke.log >> profiles/r2_pytest_gpu.txt
