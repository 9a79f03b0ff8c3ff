#!/bin/bash
for n in 128 1024; do
for cfg in "1 0" "2 0" "3 0" "4 0" "3 1"; do
  set -- $cfg
  a=$(SEA_B200_LIB=sea_codec_b200/variants/libsea_b200_prev.so python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  b=$(python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  echo "n=$n bits=$1 vbr=$2: prev=$a current=$b"
done
done
