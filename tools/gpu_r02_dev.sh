#!/bin/bash
# quick development loop on one GPU: parity suite (no variant sweep), bench at N = 1
TAG=${1:-r02d}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not gather_kernel_variants" > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu $BENCH_ARGS > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench.json"))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"])
    print(d["detail"]["phase_ms_per_step"], d["detail"]["gmres_spmv_count"], d["detail"]["iNorm"])
    print("spmv frac",d["roofline"]["frac"])
except Exception as ex:
    print("bench failed",ex); print(open("$O/${TAG}_bench.err").read()[-2000:])
PY
