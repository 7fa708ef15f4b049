#!/bin/bash
# Round 2, after the small-block SpMV families became the default: full GPU test suite, NSSOLVER and heat bench lines at
# 10M tets, the default bench line, ncu launch list of the NSSOLVER step.   gpurun --timeout 600 -- bash tools/gpu_r02_small_final.sh
TAG=${1:-r02s}
O=gpurun_out
mkdir -p $O
rm -f $O/${TAG}_parity.log
SVFSI_PARITY_LOG=$O/${TAG}_parity.log timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --solver ns > $O/${TAG}_bench_ns.json 2> $O/${TAG}_bench_ns.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --physics heat > $O/${TAG}_bench_heat.json 2> $O/${TAG}_bench_heat.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file $O/${TAG}_launches_ns.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --solver ns > $O/${TAG}_launches_ns.log 2>&1
python - <<'PY'
import json
for n in ("ns", "heat", ""):
    f = f"gpurun_out/r02s_bench{'_' + n if n else ''}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(n or "gmres", round(d["value"], 2), "it/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 2),
              "roof", d["roofline"]["kernel"][:40], round(d["roofline"]["frac"], 3), d["detail"]["phase_ms_per_step"])
    except Exception as ex:
        print(n, "failed", ex)
PY
