#!/bin/bash
# round-2 source-level captures (one GPU): encode at 128 / 1024 streams, VBR decode, 8-channel decode, stereo CBR decode
N="ncu --set full --import-source on --clock-control none --launch-skip 1 -c 1"
$N -k regex:encode_kernel -f -o gpurun_out/r02_enc_cbr3_128 python tools/enc_probe.py 128 10 3 0 > gpurun_out/ncu1.log 2>&1
$N -k regex:encode_kernel -f -o gpurun_out/r02_enc_cbr3_1024 python tools/enc_probe.py 1024 10 3 0 >> gpurun_out/ncu1.log 2>&1
PROBE_VBR=1 $N -k regex:decode_vbr_kernel -f -o gpurun_out/r02_dec_vbr3 python tools/dec_probe.py 1024 60 3 2 3 >> gpurun_out/ncu1.log 2>&1
$N -k regex:decode_mc_kernel -f -o gpurun_out/r02_dec_mc8 python tools/dec_probe.py 64 60 4 8 3 >> gpurun_out/ncu1.log 2>&1
$N -k regex:decode_unrolled_kernel -f -o gpurun_out/r02_dec_unrolled python tools/dec_probe.py 1024 60 3 2 3 >> gpurun_out/ncu1.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu1.log
