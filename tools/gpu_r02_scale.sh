#!/bin/bash
# N GPUs: multi-rank parity check (slabs + blocks) then the bench at that N
N=${1:-2}
TAG=${2:-r02n$N}
O=gpurun_out
mkdir -p $O
if [ "$3" != "nocheck" ]; then bash tools/gpu_r02_multi.sh $N $TAG | grep -c PASS; grep -h "FAIL\|rc=" $O/${TAG}_check_*.log | head; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) \
  bench.py --gpus $N --steps 5 --warmup 3 $BENCH_ARGS > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/${TAG}_bench.json") if l.startswith("{")][-1])
    print("N=$N value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"],d["comm"])
    print(d["detail"]["phase_ms_per_step"], d["detail"]["gmres_spmv_count"], d["detail"]["iNorm"])
    print("spmv frac",d["roofline"]["frac"])
except Exception as ex:
    print("bench failed",ex); print(open("$O/${TAG}_bench.err").read()[-3000:])
PY
