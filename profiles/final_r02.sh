#!/bin/bash
# Round-2 final evidence (1 GPU): gpurun --timeout 2400 -- 'bash profiles/final_r02.sh'
# tests, smoke(), the bench line of every workload (CPU arm included), the opt-in fp16-attention line, and the two ncu launch lists.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in vga_lightglue mp1_lightglue seq_superglue superpoint_only small_stop; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err
  echo "$w: $(python -c "import json;d=json.load(open('$O/r02_bench_$w.json'));print(round(d['value'],1),d['unit'],'e2e',round(d['e2e']['value'],1),'1thr',round(d['e2e'].get('single_thread',0),1),'cpu',round(d['cpu_baseline']['value'],3),d['cpu_baseline']['layout'],'frac',round(d['roofline']['frac'],3),'launches',d['gpu_launches'])" 2>&1 | tail -1)"
done
timeout 600 python bench.py --fp16-attention --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_vga_lightglue_fp16attn.json 2> $O/r02_bench_fp16.err
echo "fp16 attention: $(python -c "import json;d=json.load(open('$O/r02_bench_vga_lightglue_fp16attn.json'));print(round(d['value'],1),'frac',round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_r02_batch.csv python profiles/capture_r02_batch.py > $O/cap_batch.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_r02_misc.csv python profiles/capture_r02_misc.py > $O/cap_misc.log 2>&1
tail -1 $O/cap_batch.log; tail -1 $O/cap_misc.log
