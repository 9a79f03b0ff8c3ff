#!/bin/bash
# encode time at low stream counts under both lane mappings (SEA_B200_ENC_SPLIT): tools/split_probe.sh
for n in 128 256 296 512; do
  for vbr in 0 1; do
    for sp in 0 1; do
      echo "n=$n vbr=$vbr split=$sp: $(SEA_B200_ENC_SPLIT=$sp python tools/enc_probe.py $n 30 3.0 $vbr | tail -1)"
    done
  done
done
