#!/bin/bash
# speculative full-sky chain: the full-sky tests + deferred tests, then the headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deferred.py tests/test_gpu_intensity.py -x -q -k "fullsky or deferred or gibbs or monopole or t_cmb or tuner or statistics" > gpurun_out/r02c_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02c_tests.log
tail -4 gpurun_out/r02c_tests.log
timeout 250 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["per_kernel"])
P
