#!/bin/bash
# Round 2, small-block SpMV families: parity tests of every family, timing table at 10M tets, ncu --set full of one
# launch per shape and family.   gpurun --timeout 600 -- bash tools/gpu_r02_small.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02small_gpu.txt 2>&1
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_unstructured.py -q -m gpu -x -k "sparmul or irregular" \
  > gpurun_out/r02small_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02small_pytest.log
tail -3 gpurun_out/r02small_pytest.log
timeout 300 python tools/bench_spmv_small.py 408 > gpurun_out/r02small_timing.json 2> gpurun_out/r02small_timing.err
echo "timing rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02small_timing.json"))
print("check", {k: f"{v:.1e}" for k, v in d["check_1M_max_rel_diff_vs_mode0"].items() if v > 1e-14} or "all <= 1e-14")
for k, v in d["timing"].items():
    print(k, f"{v['ms']*1e3:.1f} us  {v['frac']:.3f}")
PY
if [ -n "$NCU_MODES" ]; then
  SMALL_MODES=$NCU_MODES timeout 300 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --clock-control none -k regex:spmv_ -c 40 -f -o gpurun_out/r02small_ncu \
    python tools/bench_spmv_small.py 408 --ncu > gpurun_out/r02small_ncu.log 2>&1
  echo "ncu rc=$?"
  ncu -i gpurun_out/r02small_ncu.ncu-rep --page raw --csv > gpurun_out/r02small_ncu_raw.csv 2>/dev/null
  rm -f gpurun_out/r02small_ncu.ncu-rep
fi
ls -la gpurun_out/r02small_*
