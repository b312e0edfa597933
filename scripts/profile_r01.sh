# Round-1 profile of the current build (1 GPU): plain run first, then launch list, then one full-set capture.
set -e
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r01b_plain.log 2>&1
tail -n 1 gpurun_out/r01b_plain.log > gpurun_out/r01b_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r01b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cg_recompute|rhs_blocks|mh_suffstat|mh_suff_chain' -s 60 -c 24 -o /tmp/r01b -f \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r01b_ncu_full.log 2>&1
ncu -i /tmp/r01b.ncu-rep --page raw --csv > gpurun_out/r01b_raw.csv
# config c4 (per-pixel beta_d, T_d) at nside 512: the per-pixel Metropolis kernel
python bench.py --config c4 --nside 512 --steps 3 --warmup 2 --no-cpu > gpurun_out/r01b_c4_plain.log 2>&1
ncu --set full --clock-control none -k regex:'mh_perpixel|chisq_kernel|rhs_blocks_kernel' -c 6 -o /tmp/r01b_c4 -f \
  python bench.py --config c4 --nside 512 --steps 1 --warmup 1 --no-cpu > gpurun_out/r01b_c4_ncu.log 2>&1
ncu -i /tmp/r01b_c4.ncu-rep --page raw --csv > gpurun_out/r01b_c4_raw.csv
# config c4 at full size on one GPU with the current build
python bench.py --config c4 --steps 3 --warmup 2 --no-cpu > gpurun_out/r01b_c4_2048_n1.log 2>&1
ls -la gpurun_out/r01b_*
