#!/bin/bash
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== cooperative"; timeout 900 python scripts/ab_run.py kmer50 kmer625 2>&1 | tail -2
echo "== plain"; RJ_KMER_PLAIN=1 timeout 900 python scripts/ab_run.py kmer50 kmer625 2>&1 | tail -2
echo "== bench headline cooperative"; RJ_BENCH_CONFIGS=0 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['per_pattern_calls']['ms_per_step'])"
echo "== bench headline plain"; RJ_KMER_PLAIN=1 RJ_BENCH_CONFIGS=0 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['per_pattern_calls']['ms_per_step'])"
