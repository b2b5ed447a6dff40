#!/bin/bash
# Round-2 evidence run (on the GPU box, under gpurun): launch lists and one `ncu --set full` capture per kernel family.
#   gpurun --timeout 2400 -- 'bash profiles/capture_r02.sh'
# Raw reports land in gpurun_out/ (scratch); `python profiles/summarize.py` turns them into the committed profiles/r02_*.txt.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_r02_batch.csv python profiles/capture_r02_batch.py > $O/cap_batch.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_r02_misc.csv python profiles/capture_r02_misc.py > $O/cap_misc.log 2>&1
full() {  # name regex skip script
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02_$1 python $4 > $O/cap_$1.log 2>&1
  echo "$1: $(tail -1 $O/cap_$1.log | cut -c1-120)"
}
full flash_ps k_flash_ps 4 profiles/capture_r02_batch.py
full gemm_ws_qkv k_gemm_ws 18 profiles/capture_r02_batch.py
full gemm_ws_ffn0 k_gemm_ws 20 profiles/capture_r02_batch.py
full ln_gelu k_lg_ln_gelu 4 profiles/capture_r02_batch.py
full nms k_nms 2 profiles/capture_r02_batch.py
full conv_tma k_conv_tma 10 profiles/capture_r02_batch.py
full rs_hyp_E k_rs_hyp_E 2 profiles/capture_r02_batch.py
full rs_refine k_rs_refine 1 profiles/capture_r02_batch.py
full sinkhorn k_sg_sinkhorn 1 profiles/capture_r02_misc.py
full sample_desc k_sample_desc 2 profiles/capture_r02_batch.py
full lg_col_argmax k_lg_col_argmax 1 profiles/capture_r02_batch.py
ls -la $O/prof_r02_*.ncu-rep | awk '{print $5, $9}'
