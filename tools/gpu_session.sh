#!/bin/bash
# One GPU-box session: parity tests, kernel-variant timings, bench (best assembly variant), ncu launch
# list + full capture.  Everything lands in gpurun_out/<tag>_*.   Usage: tools/gpu_session.sh <tag>
TAG=${1:-r01s5}
O=gpurun_out
mkdir -p $O
rm -f $O/${TAG}_variants_ok.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
export SVFSI_VARIANT_OK_FILE=$PWD/$O/${TAG}_variants_ok.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
unset SVFSI_VARIANT_OK_FILE
# the SPARMULVV quad kernel (env-selected) through every test that multiplies or solves
SVFSI_SPMV_QUAD=1 timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider \
    -k "sparmul or gmres or nssolver or time_loop or irregular or host_matrix" > $O/${TAG}_pytest_spmvquad.log 2>&1
echo "pytest(spmv quad) rc=$?" >> $O/${TAG}_pytest_spmvquad.log
timeout 400 python tools/time_asm_variants.py 408 > $O/${TAG}_asm_variants.json 2> $O/${TAG}_asm_variants.err
BEST=$(python tools/pick_asm_tune.py $O/${TAG}_asm_variants.json $O/${TAG}_variants_ok.txt 2>> $O/${TAG}_asm_variants.err || echo 40)
echo "best tune: $BEST" > $O/${TAG}_best_tune.txt
export SVFSI_ASM_TUNE=$BEST
timeout 500 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
if [ "${QUICK:-0}" = "1" ]; then tail -3 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_pytest_spmvquad.log; cat $O/${TAG}_best_tune.txt; cat $O/${TAG}_asm_variants.json; cut -c1-300 $O/${TAG}_bench.json; exit 0; fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $O/${TAG}_launches_10M.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"fluid_record|fluid_gather" -c 4 -f -o $O/${TAG}_prof_asm \
    python bench.py --nz 104 --steps 1 --warmup 1 --no-cpu > $O/${TAG}_prof_asm.log 2>&1
ncu -i $O/${TAG}_prof_asm.ncu-rep --page raw --csv > $O/${TAG}_prof_asm_raw.csv 2>/dev/null
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_best_tune.txt; cat $O/${TAG}_asm_variants.json; cut -c1-600 $O/${TAG}_bench.json
