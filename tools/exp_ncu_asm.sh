#!/bin/bash
# ncu --set full of the gather-assembly kernels (first and second generation) at 2.5M tets
O=gpurun_out
for T in 790656 1048576; do
  SVFSI_ASM_TUNE=$T timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"fluid_record|fluid_gather" -c 3 -o $O/r02_prof_asm_$T -f \
    python bench.py --nz 104 --steps 1 --warmup 1 --no-cpu > $O/r02_prof_asm_$T.log 2>&1
  tail -2 $O/r02_prof_asm_$T.log
done
ls -la $O/r02_prof_asm_*.ncu-rep
