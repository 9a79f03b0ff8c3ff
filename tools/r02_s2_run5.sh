#!/bin/bash
# session-2 run 5: whole-frame kernel with the window loaded once per cycle of phases (variant build) against the default
for lib in "" $PWD/sea_codec_b200/variants/libsea_b200_wincycle.so; do
  export SEA_B200_LIB=$lib; [ -z "$lib" ] && unset SEA_B200_LIB
  python -m pytest tests/test_gpu_round2.py tests/test_gpu_next_rows.py -m gpu -x -q -k "odd_channel or multichannel or reference_c_decoder" 2>&1 | tail -2
  for ch in 3 5 6 7; do python tools/dec_probe.py $((2048/ch)) 60 3 $ch 6 $((5120*500)); done
  python tools/dec_probe.py 682 60 5 3 6 $((5120*500))
  python tools/dec_probe.py 341 60 5 6 6 $((5120*500))
  python tools/dec_probe.py 1024 20 3 3 6
done
