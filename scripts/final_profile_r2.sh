#!/bin/bash
# Round-2 profile capture on one B200 (run through gpurun).  Every program is first run WITHOUT ncu and must
# exit 0; numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-verify --cg-iters 10"
$B > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_prof_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1
python scripts/profile_target.py 128 lagrange1 1 2 > gpurun_out/r2_prof_target.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_rowhash|k_local|k_numeric_t|k_spmv|k_cg_xr|k_cg_p|k_inc_count|k_inc_place|k_compact' -c 12 -f \
    -o gpurun_out/r2_full_assembly python scripts/profile_target.py 128 lagrange1 1 2 > gpurun_out/r2_prof_ncu2.log 2>&1
ncu -i gpurun_out/r2_full_assembly.ncu-rep --page raw --csv > gpurun_out/r2_full_assembly.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_full_assembly.csv > gpurun_out/r2_ncu_full_assembly.txt 2>&1
python scripts/peer_vs_single.py 96 3 darcy > gpurun_out/r2_prof_peer.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'kd_spmv|kd_mr_d|kd_mr_c|kd_cg_p|kd_cg_xr|k_mr_d|k_spmv' -c 14 -f \
    -o gpurun_out/r2_full_krylov python scripts/peer_vs_single.py 96 1 darcy > gpurun_out/r2_prof_ncu3.log 2>&1
ncu -i gpurun_out/r2_full_krylov.ncu-rep --page raw --csv > gpurun_out/r2_full_krylov.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_full_krylov.csv > gpurun_out/r2_ncu_full_krylov.txt 2>&1
rm -f gpurun_out/r2_full_assembly.ncu-rep gpurun_out/r2_full_krylov.ncu-rep  # keep the CSV exports only (size)
cat gpurun_out/r2_ncu_full_assembly.txt; cat gpurun_out/r2_ncu_full_krylov.txt; head -30 gpurun_out/r2_launches_summary.txt
