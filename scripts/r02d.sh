#!/bin/bash
# deferred-scalars check on one GPU: its tests, then the bench with and without it (same box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_deferred.py -x -q > gpurun_out/r02d_tests.log 2>&1
echo "deferred tests exit $?" >> gpurun_out/r02d_tests.log
tail -5 gpurun_out/r02d_tests.log
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --defer-scalars 0 > gpurun_out/r02d_bench_sync.json 2> gpurun_out/r02d_bench_sync.err
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --defer-scalars 1 > gpurun_out/r02d_bench_defer.json 2> gpurun_out/r02d_bench_defer.err
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --defer-scalars 0 > gpurun_out/r02d_bench_sync2.json 2>> gpurun_out/r02d_bench_sync.err
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --defer-scalars 1 > gpurun_out/r02d_bench_defer2.json 2>> gpurun_out/r02d_bench_defer.err
for f in sync defer sync2 defer2; do python - "$f" <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r02d_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["gpu_launches"], d["clocks"])
except Exception as e:
    print(f, "FAILED", e)
P
done
tail -3 gpurun_out/r02d_bench_defer.err
