#!/bin/bash
# one 8-GPU box: fused-exchange parity on 2 GPUs, then the weak-scaling ladder N = 1, 2, 4, 8 for C2 (halo)
# and C5 (allgather; scale 24 + log2 N, i.e. BASELINE's scale 27 at N = 8), fused exchange vs NCCL at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
run() { name=$1; N=$2; shift; shift
  if [ "$N" = 1 ]; then
    env "$@" timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 $EXTRA > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err
  else
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 $EXTRA > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err
  fi
  grep '^{' gpurun_out/scale_$name.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); c=d['config']
    print('$name', 'N=%d'%d['n_gpus'], 'ms=%.4f'%d['ms_per_step'], 'GFLOP/s=%.0f'%d['value'], c.get('exchange'), '|', c.get('exchange_impl'), '| kern', d['roofline'].get('kernel_ms', c.get('kernel_only_ms')), c.get('fused_error'), 'e2e', (d.get('e2e') or {}).get('value'))
" | tee -a gpurun_out/scale_summary.txt
}
: > gpurun_out/scale_summary.txt
export SPBLAS_B200_NO_CUSPARSE=1
EXTRA="--no-cpu-baseline"
for N in 1 2 4 8; do run c2_n$N $N SPBLAS_B200_FUSED=1; done
run c2_n8_nccl 8 SPBLAS_B200_FUSED=0
EXTRA="--workload c5"
for N in 1 2 4 8; do run c5_n$N $N SPBLAS_B200_FUSED=1; done
run c5_n8_nccl 8 SPBLAS_B200_FUSED=0
run c5_n8_mc 8 SPBLAS_B200_FUSED=1 SPBLAS_B200_MULTICAST=1
tail -3 gpurun_out/scale_*.err | tail -40
