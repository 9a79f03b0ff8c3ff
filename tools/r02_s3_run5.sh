#!/bin/bash
# session 3 run 5: uploads ahead of the lanes (own stream, one slot per group) against uploads tied to the lanes
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_next_rows.py -m gpu -x -q 2>&1 | tail -2
python tools/e2e_probe.py 512 60 > gpurun_out/r02s3_e2e_probe_ahead.txt 2> gpurun_out/r02s3_e2e_probe_ahead.err
cat gpurun_out/r02s3_e2e_probe_ahead.txt
