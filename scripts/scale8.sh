# Round-2 multi-GPU measurements on one 8-GPU box (each step under its own timeout):
#   bash scripts/scale8.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r02t}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 400 $T --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/${tag}_c2_n$n.log 2>&1
  tail -n 1 gpurun_out/${tag}_c2_n$n.log | cut -c1-200
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/${tag}_c2_n1.log 2>&1
tail -n 1 gpurun_out/${tag}_c2_n1.log | cut -c1-200
# the same without deferred scalars (A/B on the same box)
for n in 8 2; do
  timeout 400 $T --nproc-per-node $n --master-port $((29650 + n)) bench.py --gpus $n --steps 100 --warmup 5 --defer-scalars 0 > gpurun_out/${tag}_c2_sync_n$n.log 2>&1
  tail -n 1 gpurun_out/${tag}_c2_sync_n$n.log | cut -c1-200
done
timeout 300 $T --nproc-per-node 8 --master-port 29611 scripts/timeline.py --mode value --steps 3 --defer 1 --out gpurun_out/${tag}_tl_value_n8.md > gpurun_out/${tag}_tl_value_n8.log 2>&1
grep comm_probe gpurun_out/${tag}_tl_value_n8.log; tail -n 1 gpurun_out/${tag}_tl_value_n8.log
DANG_MGC_NSIDE=64 timeout 400 $T --nproc-per-node 8 --master-port 29631 tests/multi_gpu_check.py > gpurun_out/${tag}_mgc8.log 2>&1
grep -c "multi-GPU parity ok" gpurun_out/${tag}_mgc8.log
timeout 400 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k world_2 > gpurun_out/${tag}_world2.log 2>&1
tail -n 2 gpurun_out/${tag}_world2.log
timeout 900 $T --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 > gpurun_out/${tag}_c4_n8.log 2>&1
tail -n 1 gpurun_out/${tag}_c4_n8.log | cut -c1-300
