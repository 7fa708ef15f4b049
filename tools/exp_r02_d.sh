#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "gather_kernel_variants and (5242880 or 7340032 or 1048576)" 2>&1 | tail -3
for T in 790656 1048576 5242880 7340032; do
  echo "tune $T"; SVFSI_ASM_TUNE=$T timeout 300 python tools/exp_asm_l2.py 408 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['nz'], d['nEl'], ' '.join('%s=%.3fms'%(k,d[k+'_ms']) for k in ('A','B','C','ABC')))"
done
