for mb in 0 48 79; do DANG_GPU_VERBOSE=1 python bench.py --no-cpu --steps 100 --opt 11=$mb > gpurun_out/s4_l2_$mb.log 2>&1; grep "dang_gpu:" gpurun_out/s4_l2_$mb.log | head -1; python - gpurun_out/s4_l2_$mb.log $mb <<'PY'
import sys, json
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line); pk = d['roofline']['per_kernel']
        print('L2', sys.argv[2], 'MB value', d['value'], 'e2e', d['e2e']['value'], 'cg', pk['cg_pass_kernel'], 'rhs', pk['rhs_blocks_kernel']['ms'], 'scalar', pk['scalar kernels']['ms'], 'suff', pk['mh_suffstat_kernel']['ms'])
PY
done
