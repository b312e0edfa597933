# Round-1 profile of the current build (1 GPU): plain run first, then launch list, then one full-set capture.
# usage: bash scripts/profile_r01.sh <tag>   -> gpurun_out/<tag>_*
set -e
tag=${1:-r01c}
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/${tag}_plain.log 2>&1
tail -n 1 gpurun_out/${tag}_plain.log > gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cg_recompute|rhs_blocks|mh_suffstat|mh_suff_chain' -s 60 -c 24 -o /tmp/${tag} -f \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ls -la gpurun_out/${tag}_*
