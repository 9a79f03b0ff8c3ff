#!/bin/bash
# session-2 run 4: 16-warp VBR variant; ncu captures of the fixed-window VBR kernel and the 3-channel whole-frame kernel
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 2 6
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 1 6
PROBE_VBR=1 python tools/dec_probe.py 1024 60 5 2 6
export SEA_B200_LIB=$PWD/sea_codec_b200/variants/libsea_b200_vbr16.so
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 2 6
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 1 6
PROBE_VBR=1 python tools/dec_probe.py 1024 60 5 2 6
unset SEA_B200_LIB
N="ncu --set full --import-source on --clock-control none -c 1"
PROBE_VBR=1 $N --launch-skip 1 -k regex:decode_vbr_kernel -f -o gpurun_out/r02s2_dec_vbr3 python tools/dec_probe.py 1024 60 3 2 3 > gpurun_out/ncus2.log 2>&1
$N --launch-skip 1 -k regex:decode_mc_kernel -f -o gpurun_out/r02s2_dec_mc3 python tools/dec_probe.py 682 60 3 3 3 $((5120*500)) >> gpurun_out/ncus2.log 2>&1
ls -la gpurun_out/r02s2_*.ncu-rep; tail -2 gpurun_out/ncus2.log
