#!/bin/bash
# Round 2, after the convolution rebuild (conv_ps.cuh): launch list of one detection + full captures of conv1b and conv2a.
#   gpurun --timeout 1200 -- 'bash profiles/capture_r02_conv.sh'
set -u
O=gpurun_out
NCU="ncu --clock-control none"
python bench.py --workload superpoint_only --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_superpoint_only_convps.json 2> $O/bench_sp.err; tail -c 200 $O/bench_sp.err
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_r02_misc2.csv python profiles/capture_r02_misc.py > $O/cap_misc2.log 2>&1
full() {  # name regex skip script
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02_$1 python $4 > $O/cap_$1.log 2>&1
  echo "$1: $(tail -1 $O/cap_$1.log | cut -c1-120)"
}
full conv_ps_1b k_conv_ps 9 profiles/capture_r02_batch.py
full conv_ps_2a k_conv_ps 10 profiles/capture_r02_batch.py
full conv_ps_3b k_conv_ps 13 profiles/capture_r02_batch.py
