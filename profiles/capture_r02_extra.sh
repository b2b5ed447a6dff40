set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:k_assign_ps -s 1 -c 1 -f -o $O/prof_r02_assign_lg python profiles/capture_r02_batch.py > $O/cap_assign.log 2>&1; tail -1 $O/cap_assign.log | cut -c1-100
B2_FP16_ATTN=1 timeout 600 $NCU --set full --import-source on -k regex:k_flash_ps -s 4 -c 1 -f -o $O/prof_r02_flash_ps_fp16 python profiles/capture_r02_batch.py > $O/cap_flash16.log 2>&1; tail -1 $O/cap_flash16.log | cut -c1-100
