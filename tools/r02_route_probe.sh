#!/bin/bash
# decode time against job size for the two routes (SEA_B200_DEC_LATENCY): one stream of k chunks
for ch in 2 8; do
  for k in 148 888 1776 3552 7104; do
    fr=$((5120*k))
    a=$(SEA_B200_DEC_LATENCY=1 python tools/dec_probe.py 1 1 3 $ch 6 $fr | tail -1 | awk '{print $6}')
    b=$(SEA_B200_DEC_LATENCY=0 python tools/dec_probe.py 1 1 3 $ch 6 $fr | tail -1 | awk '{print $6}')
    echo "channels=$ch chunks=$k: small-job kernel $a ms   throughput kernels $b ms"
  done
done
