#!/bin/bash
# same-box A/B of the fused SpMV + halo-send kernel on 2 GPUs: 8 lanes per row (0) vs 4 lanes per row (1)
O=gpurun_out; mkdir -p $O
for FQ in 0 1; do
  SVFSI_SPMV_FUSED_QUAD=$FQ timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 \
      --master-port $((29520 + FQ)) bench.py --gpus 2 --steps 5 --warmup 3 > $O/r01n2ab_fq${FQ}.json 2> $O/r01n2ab_fq${FQ}.err
done
for FQ in 0 1; do grep -o '"value": [0-9.]*' $O/r01n2ab_fq${FQ}.json | head -1; grep -o '"spmv": [0-9.]*' $O/r01n2ab_fq${FQ}.json; done
