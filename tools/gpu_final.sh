#!/bin/bash
# Round-end GPU session (one box, one GPU): parity tests, ncu --set full of the dominant kernel
# (-> roofline.traffic), bench, ncu launch list of the bench command, ncu --set full of the assembly
# kernels, kernel-variant timings.  Everything lands in gpurun_out/<tag>_*; the summaries that are
# judged are copied into profiles/ by hand afterwards.      Usage: tools/gpu_final.sh <tag>
TAG=${1:-r01final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
export SVFSI_VARIANT_OK_FILE=$PWD/$O/${TAG}_variants_ok.txt
rm -f $SVFSI_VARIANT_OK_FILE
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
unset SVFSI_VARIANT_OK_FILE
# 1. dominant kernel, full capture at the bench workload -> DRAM traffic per launch
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"spmv_vv4" --launch-skip 8 -c 3 -f \
    -o $O/${TAG}_prof_spmv python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_prof_spmv.log 2>&1
ncu -i $O/${TAG}_prof_spmv.ncu-rep --page raw --csv > $O/${TAG}_prof_spmv_raw.csv 2>/dev/null
python tools/ncu_spmv_traffic.py $O/${TAG}_prof_spmv_raw.csv 25463369 1728025 $O/${TAG}_spmv_traffic.json \
    "ncu --set full --clock-control none -k regex:spmv_vv4 --launch-skip 8 -c 3, python bench.py --steps 1 --warmup 1 --no-cpu, 10.03M tets (gpurun_out/${TAG}_prof_spmv.ncu-rep)" \
    > /dev/null 2>> $O/${TAG}_prof_spmv.log && cp $O/${TAG}_spmv_traffic.json profiles/r01_spmv_traffic.json
# 2. the bench lines (own arm, reference arm)
timeout 500 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
# 3. launch list of the bench command
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $O/${TAG}_launches_10M.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_launches_bench.log 2>&1
# 4. assembly kernels, full capture (2.56M tets: the kernels scale linearly)
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"fluid_record|fluid_gather" -c 3 -f -o $O/${TAG}_prof_asm \
    python bench.py --nz 104 --steps 1 --warmup 1 --no-cpu > $O/${TAG}_prof_asm.log 2>&1
ncu -i $O/${TAG}_prof_asm.ncu-rep --page raw --csv > $O/${TAG}_prof_asm_raw.csv 2>/dev/null
# 5. kernel variants
ASM_TUNES="8,40,9216,57344,139264,266240,790528" timeout 300 python tools/time_asm_variants.py 408 \
    > $O/${TAG}_asm_variants.json 2> $O/${TAG}_asm_variants.err
# 6. the other two configurations (BASELINE configs[3], [4] shapes at 10M tets)
timeout 200 python bench.py --steps 5 --warmup 3 --solver ns > $O/${TAG}_bench_ns.json 2> $O/${TAG}_bench_ns.err
timeout 200 python bench.py --steps 5 --warmup 3 --physics heat > $O/${TAG}_bench_heat.json 2> $O/${TAG}_bench_heat.err
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_spmv_traffic.json; cat $O/${TAG}_asm_variants.json; cut -c1-400 $O/${TAG}_bench.json
