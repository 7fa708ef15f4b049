#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "gather_kernel_variants and (76288128 or 109842560 or 143396992 or 176951424)" 2>&1 | tail -3
for T in 42733696 76288128 109842560 143396992 176951424; do
  echo "tune $T"; SVFSI_ASM_TUNE=$T timeout 300 python tools/exp_asm_l2.py 408 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['nz'], d['nEl'], ' '.join('%s=%.3fms'%(k,d[k+'_ms']) for k in ('A','B','C','ABC')))"
done
