#!/bin/bash
# The multi-GPU measurements of a round, on ONE box with N GPUs: `tools/run_scaling.sh N [tag]` (under gpurun --gpus N).
# Writes gpurun_out/<tag>_n<N>_*.log; every bench line carries its own `verified` block.
# CFGS="c2 c3" selects the bench configs, TOOL=0 skips the single-process multi-device runs.
N=${1:-8}; TAG=${2:-r2}; OUT=gpurun_out; mkdir -p $OUT
CFGS=${CFGS:-"c2 c3 c4 c5"}; TOOL=${TOOL:-1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
[ "$N" = 1 ] && TR="python"
(nproc; free -g | head -2; nvidia-smi topo -m | head -12) > $OUT/${TAG}_n${N}_box.txt 2>&1
# the product's own multi-device entry point: one process, one table, one result
DEVS=$(seq -s, 0 $((N-1)))
if [ "$TOOL" = 1 ]; then
timeout 600 python tools/multi_gpu_run.py --devices $DEVS --queries-per-gpu 1000000 > $OUT/${TAG}_n${N}_multi_tool.log 2>&1; tail -c 1200 $OUT/${TAG}_n${N}_multi_tool.log; echo
timeout 600 python tools/multi_gpu_run.py --devices $DEVS --queries-per-gpu 1000000 --text-refs --skip-single > $OUT/${TAG}_n${N}_multi_tool_refs.log 2>&1; tail -c 600 $OUT/${TAG}_n${N}_multi_tool_refs.log; echo
timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "sharded or scattered" > $OUT/${TAG}_n${N}_pytest_multi.log 2>&1; tail -3 $OUT/${TAG}_n${N}_pytest_multi.log
fi
# one process per GPU (what the driver launches)
for cfg in $CFGS; do
  steps=8; [ $cfg = c5 ] && steps=2
  timeout 900 $TR bench.py --gpus $N --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-file-arm > $OUT/${TAG}_n${N}_bench_$cfg.log 2> $OUT/${TAG}_n${N}_bench_$cfg.err
  echo "== $cfg rc=$?"; grep '^{"metric"' $OUT/${TAG}_n${N}_bench_$cfg.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d.get('roofline') or {}
print(d['config']['config'], 'n', d['n_gpus'], d['scaling'], 'value %.1f M q/s' % (d['value']/1e6), 'ms/step %.2f' % d['ms_per_step'], 'rows/s %.2f G' % (d['hit_rows_per_s']/1e9), 'text %.0f GB/s' % d['text_gb_per_s'], '| e2e %.2f M q/s %.1f GB/s (h2d peak %.0f)' % (d['e2e']['value']/1e6, d['e2e']['text_gb_per_s'], d['e2e']['pinned_h2d_concurrent_gb_per_s']), '| tile frac', r.get('frac'), 'whole', r.get('whole_step_frac'), '| verified', d.get('verified'), '| dl', (d.get('value_with_result_download') or {}).get('value'))
" 2>&1; tail -2 $OUT/${TAG}_n${N}_bench_$cfg.err
done
