#!/bin/bash
# Round 2, call AK: smoke() as the driver runs it, then the whole GPU tier once more on the final tree.
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2ak_smoke.txt
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -5 | tee gpurun_out/r2ak_pytest.log
