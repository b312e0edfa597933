# Round-2 multi-GPU measurements on one 8-GPU box (each step under its own timeout):
#   bash scripts/scale8.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r02s}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 400 $T --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_c2_n$n.log 2>&1
  tail -n 1 gpurun_out/${tag}_c2_n$n.log | cut -c1-200
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_c2_n1.log 2>&1
timeout 300 $T --nproc-per-node 8 --master-port 29611 scripts/timeline.py --mode value --steps 3 --out gpurun_out/${tag}_tl_value_n8.md > gpurun_out/${tag}_tl_value_n8.log 2>&1
grep comm_probe gpurun_out/${tag}_tl_value_n8.log; tail -n 1 gpurun_out/${tag}_tl_value_n8.log
timeout 300 $T --nproc-per-node 8 --master-port 29612 scripts/timeline.py --mode e2e --steps 3 --out gpurun_out/${tag}_tl_e2e_n8.md > gpurun_out/${tag}_tl_e2e_n8.log 2>&1
tail -n 1 gpurun_out/${tag}_tl_e2e_n8.log
for n in 8 4; do
  timeout 200 $T --nproc-per-node $n --master-port $((29620 + n)) scripts/pcie_probe.py > gpurun_out/${tag}_pcie_n$n.log 2>&1
  tail -n 1 gpurun_out/${tag}_pcie_n$n.log | cut -c1-400
done
DANG_MGC_NSIDE=64 timeout 400 $T --nproc-per-node 8 --master-port 29631 tests/multi_gpu_check.py > gpurun_out/${tag}_mgc8.log 2>&1
grep -c "multi-GPU parity ok" gpurun_out/${tag}_mgc8.log
timeout 900 $T --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 > gpurun_out/${tag}_c4_n8.log 2>&1
tail -n 1 gpurun_out/${tag}_c4_n8.log | cut -c1-300
