#!/bin/bash
# VBR decode table replication sweep + the 8-channel kernel after the CTA-width / one-shift changes
for bits in 3 4.5 5.5 6.5; do
  for rs in 0 3 4 5; do
    echo "vbr bits=$bits RS<=$rs: $(PROBE_VBR=1 SEA_B200_VBR_RS=$rs python tools/dec_probe.py 1024 60 $bits 2 6 | tail -1)"
  done
done
echo "mono vbr3: $(PROBE_VBR=1 python tools/dec_probe.py 1024 60 3 1 6 | tail -1)"
echo "8ch cbr4 256: $(python tools/dec_probe.py 256 60 4 8 6 | tail -1)"
echo "8ch cbr4 64: $(python tools/dec_probe.py 64 60 4 8 6 | tail -1)"
echo "4ch cbr3 256: $(python tools/dec_probe.py 256 60 3 4 6 | tail -1)"
echo "6ch cbr3 256: $(python tools/dec_probe.py 256 60 3 6 6 | tail -1)"
