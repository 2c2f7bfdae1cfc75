#!/bin/bash
# bands8k (BASELINE configs[3]) on all GPUs of the box, default NCCL settings and with the collectives limited to a few CTAs
N=${1:-8}
for tag in default cta8; do
  if [ $tag = cta8 ]; then export NCCL_MAX_CTAS=8; else unset NCCL_MAX_CTAS; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload bands8k --steps 6 --warmup 2 > gpurun_out/r2_bands_n${N}_$tag.json 2> gpurun_out/r2_bands_n${N}_$tag.err
  tail -n 1 gpurun_out/r2_bands_n${N}_$tag.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', d['ms_per_step'], d['latency'], d['exposed_comm_frac'], d['strong_scaling_efficiency'])"
done
