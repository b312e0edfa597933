#!/bin/bash
# K5 one-thread-per-pixel form: parity tests, then c4 at nside 512 with the lane form (12=1) and the pixel form (12=3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "perpixel" > gpurun_out/r02k_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02k_tests.log
tail -4 gpurun_out/r02k_tests.log
for f in 1 3; do
  timeout 300 python bench.py --config c4 --nside 512 --steps 10 --warmup 3 --no-cpu --opt 12=$f > gpurun_out/r02k_c4_512_f$f.json 2> gpurun_out/r02k_c4_512_f$f.err
  python - "$f" <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r02k_c4_512_f{f}.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("form",f,"it/s",d["value"],"ms",d["ms_per_step"],{k:(v["launches"],v["ms"]) for k,v in pk.items() if "perpixel" in k or "rhs" in k}, d["config"].get("k5_fp64_fallbacks_per_proposal"))
except Exception as e:
    print(f,"FAILED",e); print(open(f"gpurun_out/r02k_c4_512_f{f}.err").read()[-1500:])
P
done
