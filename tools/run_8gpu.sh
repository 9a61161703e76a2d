#!/bin/bash
# One 8-GPU lease.  usage (GPU box): bash tools/run_8gpu.sh [tests] [bench] [config5]
#   tests   : the 8-rank parity cases of tests/test_gpu_multi.py
#   bench   : the headline bench (eigen_s N = 50000) on the 2x4 grid
#   config5 : BASELINE configs[4], eigen_sx N = 100000 on the 2x4 grid (size / solver through the environment:
#             torch.distributed.run's own parser claims an abbreviated --n)
mkdir -p gpurun_out
echo "gpus: $(nvidia-smi -L | wc -l)"
WHAT="${*:-tests bench config5}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618"
for w in $WHAT; do
  case $w in
    tests)
      timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8-" > gpurun_out/r02_multi8_tests.log 2>&1
      tail -4 gpurun_out/r02_multi8_tests.log ;;
    bench)
      timeout 500 $TR bench.py --gpus 8 --steps 5 --warmup 3 --budget-s 230 > gpurun_out/r02_bench_n8.log 2> gpurun_out/r02_bench_n8.err
      tail -c 1500 gpurun_out/r02_bench_n8.log; tail -2 gpurun_out/r02_bench_n8.err ;;
    config5)
      EIGENEXA_BENCH_N=100000 EIGENEXA_BENCH_SOLVER=sx timeout 600 $TR bench.py --gpus 8 --steps 1 --warmup 0 --cold --budget-s 330 --no-e2e \
        > gpurun_out/r02_bench_sx_n100000_8.log 2> gpurun_out/r02_bench_sx_n100000_8.err
      tail -c 1800 gpurun_out/r02_bench_sx_n100000_8.log; tail -3 gpurun_out/r02_bench_sx_n100000_8.err ;;
  esac
done
