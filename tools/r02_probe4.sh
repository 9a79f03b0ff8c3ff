#!/bin/bash
for n in 128 1024; do
  for cfg in "3 0" "5 0" "3 1"; do
    set -- $cfg
    echo "enc n=$n bits=$1 vbr=$2: $(SEA_B200_ENC_SPLIT=0 python tools/enc_probe.py $n 30 $1 $2 | tail -1)"
  done
done
