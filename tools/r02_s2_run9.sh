#!/bin/bash
# session-2 run 9: upper bound of what hiding the per-block code / scale-factor writes of the encoder could buy (probe build: wrong output)
for n in 128 1024; do for cfg in "3 0" "5 0" "3 1"; do set -- $cfg
  a=$(python tools/enc_probe.py $n 20 $1 $2 | tail -1)
  b=$(SEA_B200_LIB=$PWD/sea_codec_b200/variants/libsea_b200_nocodes.so python tools/enc_probe.py $n 20 $1 $2 | tail -1)
  echo "n=$n bits=$1 vbr=$2: default: $a"; echo "n=$n bits=$1 vbr=$2: nocodes: $b"
done; done
