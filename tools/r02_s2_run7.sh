#!/bin/bash
# session-2 run 7: sector-paired row fetch again, with the carve-out pinned to the maximum and the chosen CTA width printed
export SEA_B200_DEBUG_LAUNCH=1
for lib in "" $PWD/sea_codec_b200/variants/libsea_b200_pair.so; do
  export SEA_B200_LIB=$lib; [ -z "$lib" ] && unset SEA_B200_LIB
  for b in 3 5 6 7 8; do python tools/dec_probe.py 1024 60 $b 2 6 2>&1 | sort | uniq -c | tail -3; done
  python tools/dec_probe.py 1024 60 3 1 6 2>&1 | sort | uniq -c | tail -3
  python tools/dec_probe.py 4096 60 3 2 30 2>&1 | sort | uniq -c | tail -3
  python tools/dec_probe.py 4096 60 8 2 30 2>&1 | sort | uniq -c | tail -3
  python tools/dec_probe.py 256 60 4 8 6 2>&1 | tail -1
  python tools/dec_probe.py 682 60 3 3 6 2>&1 | tail -1
done
