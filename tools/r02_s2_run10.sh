#!/bin/bash
# session-2 run 10: encode epilogue -- mode as a constant for the CBR instances, one put_bits for codes and scale factors (A/B on one box)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "encode or golden or config" 2>&1 | tail -2
V=$PWD/sea_codec_b200/variants
for n in 128 1024; do for cfg in "1 0" "3 0" "5 0" "3 1" "5 1"; do set -- $cfg
  a=$(SEA_B200_LIB=$V/libsea_b200_encold.so python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  b=$(SEA_B200_LIB=$V/libsea_b200_encmode.so python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  c=$(python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  a2=$(SEA_B200_LIB=$V/libsea_b200_encold.so python tools/enc_probe.py $n 20 $1 $2 | tail -1 | awk '{print $3}')
  echo "n=$n bits=$1 vbr=$2: old=$a (again $a2) mode_const=$b mode_const+one_put=$c ms"
done; done
