#!/bin/bash
# experiments of one session: gather5 assembly kernels, stream SpMV shapes, NS / heat benches
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "(gather_kernel_variants and (1048576 or 3145728 or 790528)) or sparmul or nssolver or heat" 2>&1 | tail -5
for T in 790656 1048576 3145728; do
  echo "tune $T"; SVFSI_ASM_TUNE=$T timeout 300 python tools/exp_asm_l2.py 104 408 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['nz'], d['nEl'], ' '.join('%s=%.3fms'%(k,d[k+'_ms']) for k in ('A','B','C','ABC')))"
done
for S in 0 1; do
  for ARGS in "--solver ns" "--physics heat"; do
    SVFSI_SPMV_STREAM=$S timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu $ARGS 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream=$S $ARGS', 'value %.2f ms %.2f'%(d['value'],d['ms_per_step']), 'spmv frac %.3f'%d['roofline']['frac'], {k:round(v,2) for k,v in d['detail']['phase_ms_per_step'].items()}, d['detail']['gm_itr'], d['detail']['cg_itr'], d['detail']['gmres_spmv_count'])"
  done
done
