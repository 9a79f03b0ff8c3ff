#!/bin/bash
N="ncu --set full --import-source on --clock-control none --launch-skip 1 -c 1"
SEA_B200_ENC_SPLIT=0 $N -k regex:encode_kernel -f -o gpurun_out/r02_enc_cbr3_128_v2 python tools/enc_probe.py 128 10 3 0 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
