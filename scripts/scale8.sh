# 8-GPU box: multi-rank parity, strong scaling of the headline config, and config c4 at full size
set -x
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
run 8 tests/multi_gpu_check.py 2>&1 | grep -E "parity ok|Error|error" | head -10
for n in 8 4 2; do run $n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/scale_c2_n$n.log 2>&1; python profiles/show_bench.py gpurun_out/scale_c2_n$n.log | head -1; done
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/scale_c2_n1.log 2>&1; python profiles/show_bench.py gpurun_out/scale_c2_n1.log | head -1
run 8 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 > gpurun_out/scale_c4_n8.log 2>&1; python profiles/show_bench.py gpurun_out/scale_c4_n8.log
nvidia-smi topo -m | head -12
