#!/bin/bash
# 2-GPU evidence: gpurun --gpus 2 --timeout 1500 -- 'bash profiles/bench_r02_2gpu.sh'
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -3
python bench.py --scaling strong --frames 120 2> $O/r02_strong_1gpu.err | grep -v "^NCCL" > $O/r02_strong_1gpu.json
echo "strong 1 gpu: $(python -c "import json;d=json.load(open('$O/r02_strong_1gpu.json'));print(round(d['value'],1),d['strong']['phases_rank0_s'])" 2>&1 | tail -1)"
$TR bench.py --gpus 2 --steps 3 --warmup 3 2> $O/r02_bench_2gpu.err | grep -v "^NCCL" > $O/r02_bench_vga_lightglue_2gpu.json
echo "weak 2 gpu: $(python -c "import json;d=json.load(open('$O/r02_bench_vga_lightglue_2gpu.json'));print(round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['n_gpus'])" 2>&1 | tail -1)"
$TR bench.py --gpus 2 --scaling strong --frames 120 2> $O/r02_strong_2gpu.err | grep -v "^NCCL" > $O/r02_strong_2gpu.json
echo "strong 2 gpu: $(python -c "import json;d=json.load(open('$O/r02_strong_2gpu.json'));print(round(d['value'],1),d['strong'])" 2>&1 | tail -1)"
tail -3 $O/r02_strong_2gpu.err | cut -c1-300
