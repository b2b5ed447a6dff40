#!/bin/bash
# Round-2 bench lines of every workload (1 GPU): gpurun --timeout 3000 -- 'bash profiles/bench_r02.sh'
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
for w in vga_lightglue mp1_lightglue seq_superglue superpoint_only small_stop; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err
  echo "$w: $(python -c "import json;d=json.load(open('$O/r02_bench_$w.json'));print(round(d['value'],1),d['unit'],'e2e',round(d['e2e']['value'],1),'cpu',round(d['cpu_baseline']['value'],3),d['cpu_baseline']['layout'],'frac',round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
done
timeout 900 python bench.py --impl reference --steps 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
echo "reference: $(python -c "import json;d=json.load(open('$O/r02_bench_reference.json'));print(round(d['value'],3),d['cpu_baseline']['layout'],[ (t['workers'],t['threads'],round(t.get('value',0),3)) for t in d['cpu_baseline']['layouts_tried']])" 2>&1 | tail -1)"
timeout 600 python bench.py --scaling strong --frames 120 > $O/r02_strong_1gpu.json 2> $O/r02_strong_1gpu.err
echo "strong 1 gpu: $(python -c "import json;d=json.load(open('$O/r02_strong_1gpu.json'));print(round(d['value'],1),d['strong'])" 2>&1 | tail -1)"
