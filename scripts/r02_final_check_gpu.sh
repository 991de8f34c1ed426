#!/bin/bash
# r02_final_check_gpu.sh [tag]: the final tree once more on the GPU: parity tests, smoke(), the single pair.
tag=${1:-r02f}
out=gpurun_out; mkdir -p $out
( time timeout 240 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc $?" >> $out/${tag}_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
timeout 60 python scripts/time_single.py 5 > $out/${tag}_time_single_c1.log 2>&1
