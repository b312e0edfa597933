#!/bin/bash
# ncu capture of the per-pixel Metropolis kernels of c4 at nside 512: lane form (12=1) and pixel form (12=3)
mkdir -p gpurun_out
for f in 1 3; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mh_perpixel' -s 2 -c 2 -o /tmp/r02k_f$f -f \
    python bench.py --config c4 --nside 512 --steps 1 --warmup 3 --no-cpu --opt 12=$f > gpurun_out/r02k_ncu_f$f.log 2>&1
  ncu -i /tmp/r02k_f$f.ncu-rep --page raw --csv > gpurun_out/r02k_f${f}_raw.csv
  ncu -i /tmp/r02k_f$f.ncu-rep --page source --csv > gpurun_out/r02k_f${f}_source.csv 2>/dev/null || true
done
ls -la gpurun_out/r02k_f*
