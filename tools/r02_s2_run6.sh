#!/bin/bash
# session-2 run 6: sector-paired row fetch of the unrolled kernel (variant build) against the default
for lib in "" $PWD/sea_codec_b200/variants/libsea_b200_pair.so; do
  export SEA_B200_LIB=$lib; [ -z "$lib" ] && unset SEA_B200_LIB
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -2
  for b in 1 3 4 5 6 7 8; do python tools/dec_probe.py 1024 60 $b 2 6; done
  python tools/dec_probe.py 1024 60 3 1 6
  python tools/dec_probe.py 1024 60 8 1 6
  python tools/dec_probe.py 4096 60 3 2 30
done
