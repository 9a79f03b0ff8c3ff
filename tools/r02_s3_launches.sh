#!/bin/bash
# session 3: launch list of the final bench command (every kernel, per-launch gpu__time_duration), after the same command ran clean
python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/r02s3_bench_plain.json 2> gpurun_out/r02s3_bench_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r02s3_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/r02s3_ncu_bench.log 2>&1
echo ncu rc=$?; wc -l gpurun_out/r02s3_launches_bench.csv
