#!/bin/bash
# tools/scale.sh N  -- the driver's launch line for N GPUs (one rank per GPU), short variant of the bench
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
  --no-cpu-baseline --no-fp32 --no-parity-sample --no-stage-profile --no-extras
