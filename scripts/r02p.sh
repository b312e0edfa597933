#!/bin/bash
# pipe-bound kernels: plain bench of c3 (nside 256) and an ncu capture of its per-pixel kernels; c4 (nside 512) capture of the final K5
mkdir -p gpurun_out
timeout 200 python bench.py --config c3 --nside 256 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02p_c3_256.json 2> gpurun_out/r02p_c3_256.err
echo "c3 bench rc=$?"; tail -c 900 gpurun_out/r02p_c3_256.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mh_perpixel' -s 2 -c 2 -o /tmp/r02p_c3 -f \
  python bench.py --config c3 --nside 256 --steps 1 --warmup 3 --no-cpu > gpurun_out/r02p_ncu_c3.log 2>&1
ncu -i /tmp/r02p_c3.ncu-rep --page raw --csv > gpurun_out/r02p_c3_raw.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mh_perpixel' -s 2 -c 2 -o /tmp/r02p_c4 -f \
  python bench.py --config c4 --nside 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/r02p_ncu_c4.log 2>&1
ncu -i /tmp/r02p_c4.ncu-rep --page raw --csv > gpurun_out/r02p_c4_raw.csv
ls -la gpurun_out/r02p_*
