#!/bin/bash
# session-2 run 1: the generalised whole-frame kernel (3..8 channels, any scale_factor_bits <= 6)
python -m pytest tests/test_gpu_round2.py tests/test_gpu_next_rows.py -m gpu -x -q -k "odd_channel or multichannel" 2>&1 | tail -5
for ch in 3 5 7 4 6 8; do python tools/dec_probe.py $((2048/ch)) 60 3 $ch 6 $((5120*500)); done
PROBE_SFB=3 python tools/dec_probe.py 256 60 4 8 6 $((5120*500))
PROBE_SFB=5 python tools/dec_probe.py 682 60 3 3 6 $((5120*500))
python tools/dec_probe.py 1024 60 3 2 6
PROBE_SFB=3 python tools/dec_probe.py 1024 60 3 2 6
PROBE_SFB=5 python tools/dec_probe.py 1024 60 3 2 6
