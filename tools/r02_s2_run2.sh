#!/bin/bash
# session-2 run 2: S-templated unrolled kernel, narrower CTAs for short grids, VBR kernel at other scale_factor_bits, cp.async.ca variant
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/dec_probe.py 1024 60 3 2 6
PROBE_SFB=3 python tools/dec_probe.py 1024 60 3 2 6
PROBE_SFB=5 python tools/dec_probe.py 1024 60 3 2 6
PROBE_SFB=6 python tools/dec_probe.py 1024 60 3 2 6
python tools/dec_probe.py 1024 20 3 2 6
python tools/dec_probe.py 1024 20 5 2 6
python tools/dec_probe.py 512 20 3 1 6
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 2 6
PROBE_VBR=1 PROBE_SFB=5 python tools/dec_probe.py 1024 60 3 2 6
PROBE_VBR=1 PROBE_SFB=3 python tools/dec_probe.py 1024 60 3 1 6
export SEA_B200_LIB=$PWD/sea_codec_b200/variants/libsea_b200_ca.so
python tools/dec_probe.py 1024 60 3 2 6
python tools/dec_probe.py 4096 60 3 2 30
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 2 6
python tools/dec_probe.py 256 60 4 8 6
unset SEA_B200_LIB
python tools/dec_probe.py 4096 60 3 2 30
