#!/bin/bash
# session-2 run 8: pair-fetch as a launch-time choice (full-width CTAs, default scale_factor_bits), default carve-out again
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export SEA_B200_DEBUG_LAUNCH=1
for b in 3 5 8; do python tools/dec_probe.py 1024 60 $b 2 6 2>&1 | sort | uniq -c | tail -2; done
python tools/dec_probe.py 1024 60 3 1 6 2>&1 | sort | uniq -c | tail -2
for b in 1 3 4 6 8; do python tools/dec_probe.py 4096 60 $b 2 8 2>&1 | sort | uniq -c | tail -2;  SEA_B200_PAIRFETCH=0 python tools/dec_probe.py 4096 60 $b 2 8 2>&1 | sort | uniq -c | tail -2; done
python tools/dec_probe.py 4096 60 3 2 30 2>&1 | tail -1
SEA_B200_PAIRFETCH=0 python tools/dec_probe.py 4096 60 3 2 30 2>&1 | tail -1
python tools/dec_probe.py 4096 60 3 1 8 2>&1 | tail -1
SEA_B200_PAIRFETCH=0 python tools/dec_probe.py 4096 60 3 1 8 2>&1 | tail -1
