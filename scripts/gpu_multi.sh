#!/bin/bash
# two GPUs: fused-exchange parity, then C2 / C5 with the fused and the NCCL exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 $EXTRA > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err
  grep '^{' gpurun_out/multi_$name.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); c=d['config']
    print('$name', 'N=%d'%d['n_gpus'], 'ms=%.4f'%d['ms_per_step'], 'GFLOP/s=%.0f'%d['value'], c.get('exchange'), '|', c.get('exchange_impl'), '| kern', d['roofline'].get('kernel_ms', c.get('kernel_only_ms')), c.get('fused_error'))
"
}
EXTRA="--no-e2e --no-cpu-baseline"
run c2_fused SPBLAS_B200_FUSED=1
run c2_nccl SPBLAS_B200_FUSED=0
EXTRA="--workload c5"
run c5_fused SPBLAS_B200_FUSED=1
run c5_mc SPBLAS_B200_FUSED=1 SPBLAS_B200_MULTICAST=1
run c5_nccl SPBLAS_B200_FUSED=0
tail -3 gpurun_out/multi_*.err | tail -30
