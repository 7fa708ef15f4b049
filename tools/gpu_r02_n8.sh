#!/bin/bash
# 8 GPUs, one session: multi-rank parity (slabs + blocks), strong-scaling bench, heat (C4), weak NS 40M (C5)
N=${1:-8}
TAG=${2:-r02n8}
O=gpurun_out
mkdir -p $O
bash tools/gpu_r02_multi.sh $N $TAG > $O/${TAG}_multi.out 2>&1
grep -h -c PASS $O/${TAG}_check_*.log; grep -h "FAIL\|rc=\|MULTI" $O/${TAG}_check_*.log | head
run() {  # name, args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
    bench.py --gpus $N --steps 5 --warmup 3 $2 > $O/${TAG}_bench_$1.json 2> $O/${TAG}_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/${TAG}_bench_$1.json") if l.startswith("{")][-1])
    print("$1 N=$N value %.2f ms %.3f e2e %.2f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]), d["comm"])
    print("  ", {k:round(v,3) for k,v in d["detail"]["phase_ms_per_step"].items()}, d["detail"]["gmres_spmv_count"], d["detail"]["gm_itr"], d["detail"]["cg_itr"], d["detail"]["iNorm"])
    print("   spmv frac %.3f"%d["roofline"]["frac"], d["clocks"]["reasons"])
except Exception as ex:
    print("$1 bench failed",ex); print(open("$O/${TAG}_bench_$1.err").read()[-2000:])
PY
}
run strong ""
if [ "$3" = "all" ]; then
run heat "--physics heat"
run weakns "--scaling weak --solver ns"
run weakgmres "--scaling weak"
fi
