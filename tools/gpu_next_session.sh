#!/bin/bash
# First GPU call of the next round (one GPU, ~3 min): everything that was written after the round-1 GPU
# budget was spent and is therefore still unverified on a B200, then the regular suite and a bench line.
#   1. tests/test_gpu_rcr.py (device-resident RCR loop, guarded by SVFSI_RUN_UNVERIFIED)
#   2. the full -m gpu suite (smoke() now also assembles with the default gather variant)
#   3. bench.py at N = 1
# Then, on 4 and 8 GPUs (separate calls): the strong-scaling line with the final kernels
#   gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
#       --master-port 29540 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json'
TAG=${1:-r02s1}
O=gpurun_out
mkdir -p $O
SVFSI_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_rcr.py -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest_rcr.log 2>&1
echo "rcr rc=$?" >> $O/${TAG}_pytest_rcr.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -3 $O/${TAG}_pytest_rcr.log; tail -3 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log; cut -c1-300 $O/${TAG}_bench.json
