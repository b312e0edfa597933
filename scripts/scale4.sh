set -x
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
run 4 tests/multi_gpu_check.py > gpurun_out/mgc4.log 2>&1; grep -E "parity ok|rror" gpurun_out/mgc4.log | head -6
for n in 4 2; do run $n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/ll_c2_n$n.log 2>&1; python profiles/show_bench.py gpurun_out/ll_c2_n$n.log | head -3; done
