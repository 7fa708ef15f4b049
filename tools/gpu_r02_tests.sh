#!/bin/bash
# one GPU: the whole -m gpu suite with every measured parity error logged, then smoke and a bench line
TAG=${1:-r02t}
O=gpurun_out
mkdir -p $O
rm -f $O/${TAG}_parity.log
SVFSI_PARITY_LOG=$O/${TAG}_parity.log timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
if [ "$2" != "nobench" ]; then
timeout 400 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c1-300 $O/${TAG}_bench.json
fi
tail -5 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log
