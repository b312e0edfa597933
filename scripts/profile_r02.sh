# Round-2 profile of the current build (1 GPU): plain run first, then the launch list, then one full-set capture of
# the three streaming kernels of the headline config.   usage: bash scripts/profile_r02.sh <tag>  -> gpurun_out/<tag>_*
set -e
tag=${1:-r02a}
python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/${tag}_plain.log 2>&1
tail -n 1 gpurun_out/${tag}_plain.log > gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cg_solve_kernel|rhs_blocks_ring|mh_suffstat_ring|mh_suff_chain' -s 8 -c 8 -o /tmp/${tag} -f \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i /tmp/${tag}.ncu-rep --page details --csv > gpurun_out/${tag}_details.csv 2>/dev/null || true
cp /tmp/${tag}.ncu-rep gpurun_out/${tag}.ncu-rep
ls -la gpurun_out/${tag}_*
