#!/bin/bash
# 2-GPU check of the final defaults (+ the fused quad SpMV kernel): parity script, then a short bench
O=gpurun_out; TAG=${1:-r01n2}; mkdir -p $O
export SVFSI_SPMV_FUSED_QUAD=${FQ:-1}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
    tests/multi_gpu_check.py > $O/${TAG}_check.log 2>&1
echo "check rc=$?" >> $O/${TAG}_check.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
grep -E "FAIL|OK|rc=" $O/${TAG}_check.log | tail -8; cut -c1-300 $O/${TAG}_bench.json
