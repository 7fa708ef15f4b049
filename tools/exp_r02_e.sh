#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_unstructured.py tests/test_gpu_c2.py -m gpu -q -p no:cacheprovider -x -k "not gather_kernel_variants or 790" 2>&1 | tail -3
for T in 790656; do
  echo "tune $T"; SVFSI_ASM_TUNE=$T timeout 300 python tools/exp_asm_l2.py 408 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['nz'], d['nEl'], ' '.join('%s=%.3fms'%(k,d[k+'_ms']) for k in ('A','B','C','ABC')))"
done
