# one full-set capture of the per-pixel Metropolis kernel at c4 / nside 512 (run under gpurun, 1 GPU)
set -e
python bench.py --config c4 --nside 512 --steps 2 --warmup 1 --no-cpu > gpurun_out/k5_plain.log 2>&1
python profiles/show_bench.py gpurun_out/k5_plain.log
ncu --set full --clock-control none --import-source on -k regex:mh_perpixel -c 2 -o /tmp/k5prof -f \
  python bench.py --config c4 --nside 512 --steps 1 --warmup 1 --no-cpu > gpurun_out/k5_ncu.log 2>&1
ncu -i /tmp/k5prof.ncu-rep --page raw --csv > gpurun_out/k5_raw.csv
ncu -i /tmp/k5prof.ncu-rep --page source --csv > gpurun_out/k5_source.csv 2>/dev/null || true
ls -la gpurun_out/k5_raw.csv gpurun_out/k5_source.csv
