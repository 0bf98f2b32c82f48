#!/bin/bash
# 4 GPUs: C2 fused vs NCCL after the exact per-tile scatter check, C5 with the automatic exchange choice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=4
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 $EXTRA > gpurun_out/scale4_$name.json 2> gpurun_out/scale4_$name.err
  grep '^{' gpurun_out/scale4_$name.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); c=d['config']
    print('$name', 'N=%d'%d['n_gpus'], 'ms=%.4f'%d['ms_per_step'], 'GFLOP/s=%.0f'%d['value'], c.get('exchange'), '|', c.get('exchange_impl'), '| kern', d['roofline'].get('kernel_ms', c.get('kernel_only_ms')), c.get('fused_error'))
" | tee -a gpurun_out/scale4_summary.txt
}
: > gpurun_out/scale4_summary.txt
export SPBLAS_B200_NO_CUSPARSE=1
EXTRA="--no-cpu-baseline --no-e2e"
run c2_fused SPBLAS_B200_FUSED=1
run c2_nccl SPBLAS_B200_FUSED=0
EXTRA="--workload c5"
run c5_auto SPBLAS_B200_FUSED=1
for f in gpurun_out/scale4_*.err; do tail -n 2 $f; done | tail -12
