#!/bin/bash
# last GPU call of round 2: launch list of the retrieval front + a sanity pass over the shipped library
O=gpurun_out
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file $O/launches_r02_retrieval.csv python profiles/capture_r02_retrieval.py > $O/cap_retrieval.log 2>&1
tail -1 $O/cap_retrieval.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests/test_lightglue_gpu.py tests/test_superpoint_gpu.py tests/test_netvlad_gpu.py tests/test_retriever_gpu.py -q 2>&1 | tail -2
