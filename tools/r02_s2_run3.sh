#!/bin/bash
# session-2 run 3: narrow-window VBR form, per-cycle ring top-up of the whole-frame kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for b in 2 3 4 5 6; do PROBE_VBR=1 python tools/dec_probe.py 1024 60 $b 2 6; done
PROBE_VBR=1 SEA_B200_VBR_KF=0 python tools/dec_probe.py 1024 60 3 2 6
PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 1 6
PROBE_VBR=1 SEA_B200_VBR_KF=0 python tools/dec_probe.py 1024 60 3 1 6
PROBE_VBR=1 PROBE_SFB=5 python tools/dec_probe.py 1024 60 3 2 6
for ch in 3 5 6 7 8; do python tools/dec_probe.py $((2048/ch)) 60 3 $ch 6 $((5120*500)); done
python tools/dec_probe.py 256 60 4 8 6 $((5120*500))
