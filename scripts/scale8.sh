# 8-GPU box: multi-rank parity, strong scaling of the headline config, and config c4 at full size
set -x
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
run 8 tests/multi_gpu_check.py > gpurun_out/mgc8.log 2>&1; grep -c "parity ok" gpurun_out/mgc8.log
run 8 bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/scale2_c2_n8.log 2>&1; python profiles/show_bench.py gpurun_out/scale2_c2_n8.log | head -3
run 8 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 > gpurun_out/scale2_c4_n8.log 2>&1; python profiles/show_bench.py gpurun_out/scale2_c4_n8.log | head -7
