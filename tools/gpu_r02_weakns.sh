#!/bin/bash
# N GPUs: config C5 (FSILS NSSOLVER, 5M tets per GPU, weak scaling) with the small-block SpMV families of round 2
N=${1:-8}
O=gpurun_out
mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --scaling weak --solver ns > $O/r02s_bench_weakns_n$N.json 2> $O/r02s_bench_weakns_n$N.err
echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$O/r02s_bench_weakns_n$N.json") if l.startswith("{")][-1])
print("value %.2f ms %.3f e2e %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["comm"], "frac %.3f"%d["roofline"]["frac"])
print({k:round(v,3) for k,v in d["detail"]["phase_ms_per_step"].items()})
PY
