set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/final_pytest.log
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python scripts/profile_target.py 128 lagrange1 2 3 > gpurun_out/final_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file gpurun_out/final_launches.csv python scripts/profile_target.py 128 lagrange1 2 3 > gpurun_out/final_ncu1.log 2>&1
python scripts/space_times.py 96 > gpurun_out/final_space96.log 2>&1
cat gpurun_out/final_pytest.log; cut -c1-300 gpurun_out/final_bench_n1.json; cut -c1-150 gpurun_out/final_space96.log
