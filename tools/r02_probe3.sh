#!/bin/bash
for n in 128 1024 4096; do
  for cfg in "3 0" "5 0" "8 0" "3 1" "5 1"; do
    set -- $cfg
    secs=30; [ $n = 4096 ] && secs=10
    echo "enc n=$n bits=$1 vbr=$2: $(python tools/enc_probe.py $n $secs $1 $2 | tail -1)"
  done
done
