"""One lock-step batch of the bench workload for ncu: 9 detections, ONE batch of 8 LightGlue pairs (640x480 frames, 5000 keypoints,
9 layers), 8 verifications.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_batch.csv python profiles/capture_r02_batch.py
  ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -s 20 -c 1 -o gpurun_out/prof_gemm python profiles/capture_r02_batch.py
"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gtsfm_b200 import synthetic as syn  # noqa: E402
from gtsfm_b200.pipeline import DeviceFrontEnd  # noqa: E402

fe = DeviceFrontEnd(syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "bench"), max_keypoints=5000,
                    fp16_attention=os.environ.get("B2_FP16_ATTN") == "1")  # B2_FP16_ATTN=1: capture the opt-in fp16 attention
frames, cal = syn.synthetic_sequence(9, 480, 640)
feats = [fe.detect(torch.from_numpy(f).cuda()) for f in frames]
res = fe.match_batch([(feats[i], feats[8]) for i in range(8)])
for i, (m, _) in enumerate(res):
    fe.verify(feats[i], feats[8], m, cal, cal, 4.0)
torch.cuda.synchronize()
print("launches", fe.ctx.launch_count(), "matches", [int(m.shape[0]) for m, _ in res])
