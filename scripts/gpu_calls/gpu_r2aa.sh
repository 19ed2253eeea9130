#!/bin/bash
# Round 2, call AA (2 GPUs): the stitch without system-scope fences.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== stitch tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "device_stitch or multi_gpu or fused_pattern_set or kmer" 2>&1 | tail -6 | tee gpurun_out/r2aa_pytest.log
for i in 1 2; do
echo "== bench --gpus 2 (headline only) run $i"
RJ_BENCH_CONFIGS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2981$i bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2aa_bench2.err | tee gpurun_out/r2aa_bench_n2_$i.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stitch_separate_launch'], d['stitch_cascades'], d['parallel_parity']['fused_totals_equal_reference'])"
done
echo "== bench N=1 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2aa_bench1.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
