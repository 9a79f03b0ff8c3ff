#!/bin/bash
# Round-2 final captures (one GPU).  Every command below first ran to completion WITHOUT ncu in the same gpurun call (bench + probes).
set -x
python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/r02_final_bench_plain.json 2> gpurun_out/r02_final_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/r02_ncu_bench.log 2>&1
N="ncu --set full --import-source on --clock-control none -c 1"
$N --launch-skip 3 -k regex:decode_unrolled_kernel -f -o gpurun_out/r02f_dec_unrolled_4096 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-encode --skip-e2e > gpurun_out/ncuf.log 2>&1
$N --launch-skip 1 -k regex:encode_kernel -f -o gpurun_out/r02f_enc_cbr3_1024 python tools/enc_probe.py 1024 10 3 0 >> gpurun_out/ncuf.log 2>&1
$N --launch-skip 1 -k regex:encode_kernel -f -o gpurun_out/r02f_enc_vbr3_1024 python tools/enc_probe.py 1024 10 3 1 >> gpurun_out/ncuf.log 2>&1
$N --launch-skip 1 -k regex:decode_mc_kernel -f -o gpurun_out/r02f_dec_mc8_256 python tools/dec_probe.py 256 60 4 8 3 >> gpurun_out/ncuf.log 2>&1
PROBE_VBR=1 $N --launch-skip 1 -k regex:decode_vbr_kernel -f -o gpurun_out/r02f_dec_vbr3 python tools/dec_probe.py 1024 60 3 2 3 >> gpurun_out/ncuf.log 2>&1
$N --launch-skip 1 -k regex:decode_staged_kernel -f -o gpurun_out/r02f_dec_staged_3ch python tools/dec_probe.py 1024 20 3 3 3 >> gpurun_out/ncuf.log 2>&1
$N --launch-skip 5 -k regex:decode_latency_kernel -f -o gpurun_out/r02f_dec_latency python tools/latency_probe.py 2 10 >> gpurun_out/ncuf.log 2>&1
ls -la gpurun_out/r02f_*.ncu-rep gpurun_out/r02_launches_bench.csv; tail -3 gpurun_out/ncuf.log
