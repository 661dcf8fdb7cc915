#!/bin/bash
# Round-2 final evidence after the k_rowins change, one B200 (run through gpurun): GPU tests, smoke, the default
# bench line, the ncu launch list of the same command and one `ncu --set full` pass over the assembly kernels.
# Every program is first run WITHOUT ncu and must exit 0; numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || exit 1
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-verify --cg-iters 10"
$B > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_prof_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1
python scripts/profile_target.py 128 lagrange1 1 2 > gpurun_out/r2_prof_target.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_rowins|k_rowhash|k_local|k_numeric_t|k_inc_count|k_inc_place|k_compact|k_pack_cells' -c 10 -f \
    -o gpurun_out/r2_full_assembly python scripts/profile_target.py 128 lagrange1 1 2 > gpurun_out/r2_prof_ncu2.log 2>&1
ncu -i gpurun_out/r2_full_assembly.ncu-rep --page raw --csv > gpurun_out/r2_full_assembly.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_full_assembly.csv > gpurun_out/r2_ncu_full_assembly.txt 2>&1
rm -f gpurun_out/r2_full_assembly.ncu-rep
cat gpurun_out/r2_ncu_full_assembly.txt; head -24 gpurun_out/r2_launches_summary.txt
