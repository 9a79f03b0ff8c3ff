#!/bin/bash
# Round-2 third session: re-verify the restored tree on a fresh box (GPU tests, smoke, the default bench run).
set -x
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/r02s3_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02s3_smoke.txt
( time python bench.py > gpurun_out/r02s3_bench.json 2> gpurun_out/r02s3_bench.err ) 2>&1 | tail -3 | tee gpurun_out/r02s3_bench_time.txt
head -c 600 gpurun_out/r02s3_bench.json
