#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): tests, smoke, both bench arms, then the
# ncu launch list and one --set full capture of the same fixed program.  Numbers printed under ncu
# are never bench values.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python scripts/profile_target.py 128 lagrange1 2 3 > gpurun_out/final_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file gpurun_out/final_launches.csv python scripts/profile_target.py 128 lagrange1 2 3 > gpurun_out/final_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_numeric_t|k_local|k_spmv' -c 4 -f -o gpurun_out/final_full \
    python scripts/profile_target.py 128 lagrange1 1 1 > gpurun_out/final_ncu2.log 2>&1
tail -2 gpurun_out/final_pytest.log; cat gpurun_out/final_smoke.log | tail -1; cat gpurun_out/final_bench_n1.json | cut -c1-400
