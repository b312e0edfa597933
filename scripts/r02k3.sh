#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3 4 5 6; do
  DANG_K5P_VARIANT=$v timeout 300 python bench.py --config c4 --nside 512 --steps 10 --warmup 3 --no-cpu --opt 12=3 > gpurun_out/r02k3_v$v.json 2> gpurun_out/r02k3_v$v.err
  python - "$v" <<'P'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r02k3_v{v}.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("variant",v,"it/s",d["value"],{k:(x["launches"],x["ms"]) for k,x in pk.items() if "perpixel" in k})
except Exception as e:
    print(v,"FAILED",e)
P
done
