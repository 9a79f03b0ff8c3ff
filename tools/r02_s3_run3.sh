#!/bin/bash
# session 3 run 3: short first group; the context on torch's (legacy default) stream against its own streams
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "pipelined or owned" 2>&1 | tail -2
python tools/e2e_probe.py 512 60 > gpurun_out/r02s3_e2e_probe_b.txt 2> gpurun_out/r02s3_e2e_probe_b.err
python tools/e2e_probe.py 512 60 own > gpurun_out/r02s3_e2e_probe_own.txt 2> gpurun_out/r02s3_e2e_probe_own.err
grep -h "control\|group=96" gpurun_out/r02s3_e2e_probe_b.txt gpurun_out/r02s3_e2e_probe_own.txt
