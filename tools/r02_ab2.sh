#!/bin/bash
for cfg in "3 0" "4 0" "5 0"; do
  set -- $cfg
  line="n=128 bits=$1:"
  for v in prev condred rtf oldargmin; do
    t=$(SEA_B200_ENC_SPLIT=0 SEA_B200_LIB=sea_codec_b200/variants/libsea_b200_$v.so python tools/enc_probe.py 128 20 $1 $2 | tail -1 | awk '{print $3}')
    line="$line $v=$t"
  done
  t=$(SEA_B200_ENC_SPLIT=0 python tools/enc_probe.py 128 20 $1 $2 | tail -1 | awk '{print $3}')
  echo "$line current=$t"
done
