#!/bin/bash
# session 3 run 2: deferred error words in the pipelined host-buffer decode -- the new test, then the e2e probe (switch A/B, group sizes, trace)
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/e2e_probe.py 512 60 > gpurun_out/r02s3_e2e_probe.txt 2> gpurun_out/r02s3_e2e_probe.err
tail -30 gpurun_out/r02s3_e2e_probe.txt
