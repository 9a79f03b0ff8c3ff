#!/bin/bash
# Round-2 (second half) final captures on one GPU.  Every command first ran to completion WITHOUT ncu in this same call.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/r02s2_bench.json 2> gpurun_out/r02s2_bench.err ) 2>&1 | tail -3
( time python bench.py --impl reference > gpurun_out/r02s2_bench_ref.json 2> gpurun_out/r02s2_bench_ref.err ) 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/r02s2_bench_plain.json 2> gpurun_out/r02s2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02s2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/r02s2_ncu_bench.log 2>&1
N="ncu --set full --import-source on --clock-control none -c 1"
$N --launch-skip 3 -k regex:decode_unrolled_kernel -f -o gpurun_out/r02s2_dec_unrolled_4096 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-encode --skip-e2e > gpurun_out/ncus2f.log 2>&1
$N --launch-skip 1 -k regex:decode_mc_kernel -f -o gpurun_out/r02s2_dec_mc8_256 python tools/dec_probe.py 256 60 4 8 3 >> gpurun_out/ncus2f.log 2>&1
ls -la gpurun_out/r02s2_*; tail -3 gpurun_out/ncus2f.log
