#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not gather_kernel_variants" 2>&1 | tail -3
for NAME in strong heat weakns weakgmres; do
  case $NAME in strong) A="";; heat) A="--physics heat";; weakns) A="--scaling weak --solver ns";; weakgmres) A="--scaling weak";; esac
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu $A > $O/r02n1b_bench_$NAME.json 2> $O/r02n1b_bench_$NAME.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r02n1b_bench_$NAME.json") if l.startswith("{")][-1])
    print("$NAME N=1 value %.2f ms %.3f e2e %.2f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]))
    print("  ", {k:round(v,3) for k,v in d["detail"]["phase_ms_per_step"].items()}, d["detail"]["gmres_spmv_count"], d["detail"]["gm_itr"], d["detail"]["cg_itr"], d["detail"]["iNorm"])
    print("   spmv frac %.3f"%d["roofline"]["frac"], d["clocks"]["reasons"], d["detail"]["nEl_rank0"])
except Exception as ex:
    print("$NAME bench failed",ex); print(open("$O/r02n1b_bench_$NAME.err").read()[-2000:])
PY
done
