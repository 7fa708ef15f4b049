#!/bin/bash
# Round-2 final session on ONE GPU: everything profiles/r02_* quotes.
TAG=${1:-r02final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
rm -f $O/${TAG}_parity.log
SVFSI_PARITY_LOG=$O/${TAG}_parity.log timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --nx 40 --nz 104 > $O/${TAG}_bench_c2_1M.json 2> $O/${TAG}_bench_c2_1M.err
timeout 300 python bench.py --steps 5 --warmup 3 --solver ns > $O/${TAG}_bench_ns.json 2> $O/${TAG}_bench_ns.err
timeout 300 python bench.py --steps 5 --warmup 3 --physics heat > $O/${TAG}_bench_heat.json 2> $O/${TAG}_bench_heat.err
timeout 300 python bench.py --steps 5 --warmup 3 --scaling weak --solver ns > $O/${TAG}_bench_weakns_n1.json 2> $O/${TAG}_bench_weakns_n1.err
timeout 300 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu > $O/${TAG}_bench_weakgmres_n1.json 2> $O/${TAG}_bench_weakgmres_n1.err
# ncu: launch list of one timed step (skip the warm-up step's launches), then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $O/${TAG}_launches_10M.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"spmv_vv4_quad_kernel|multidot_fused|multi_axpy_scale|spmv_vv4_quad_scale" -s 20 -c 4 -o $O/${TAG}_prof_la -f \
  python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_prof_la.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fluid_record|fluid_gather" -c 3 -o $O/${TAG}_prof_asm -f \
  python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_prof_asm.log 2>&1
tail -3 $O/${TAG}_pytest.log; tail -1 $O/${TAG}_smoke.log; cut -c1-200 $O/${TAG}_bench.json; ls -la $O/${TAG}_prof_*.ncu-rep
