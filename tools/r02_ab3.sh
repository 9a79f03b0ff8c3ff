#!/bin/bash
for n in 128 1024; do
for cfg in "1 0" "3 0" "4 0" "5 0" "8 0" "3 1" "5 1" "7 1"; do
  set -- $cfg
  line="n=$n bits=$1 vbr=$2:"
  for v in prev vbrrt; do
    t=$(SEA_B200_ENC_SPLIT=0 SEA_B200_LIB=sea_codec_b200/variants/libsea_b200_$v.so python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
    line="$line $v=$t"
  done
  t=$(SEA_B200_ENC_SPLIT=0 python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  echo "$line current=$t"
done
done
