#!/bin/bash
# same-box A/B of two builds: tools/r02_ab.sh <variant-name> [streams...]
V=sea_codec_b200/variants/libsea_b200_$1.so; shift
for n in "$@"; do
  for cfg in "3 0" "4 0" "5 0" "8 0" "3 1" "5 1"; do
    set -- $cfg
    a=$(SEA_B200_ENC_SPLIT=0 SEA_B200_LIB=$V python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
    b=$(SEA_B200_ENC_SPLIT=0 python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
    echo "n=$n bits=$1 vbr=$2: variant $a ms   current $b ms"
  done
done
