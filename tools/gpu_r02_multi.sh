#!/bin/bash
# N GPUs (gpurun --gpus N): multi-rank parity against the oracle on the slab AND the block (quadrant)
# partition, every measured error appended to gpurun_out/<tag>_parity_multi.log
N=${1:-4}
TAG=${2:-r02m$N}
O=gpurun_out
mkdir -p $O
for PART in slabs blocks; do
  if [ $PART = blocks ] && [ $((N % 4)) -ne 0 ]; then continue; fi
  SVFSI_PARTITION=$PART SVFSI_PARITY_LOG=$O/${TAG}_parity_multi.log timeout 600 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) tests/multi_gpu_check.py \
    > $O/${TAG}_check_$PART.log 2>&1
  echo "rc=$?" >> $O/${TAG}_check_$PART.log
  tail -4 $O/${TAG}_check_$PART.log
done
