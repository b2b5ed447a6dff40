"""A miniature bench step for ncu: 2 detections + 2 pair matches + 2 verifications on the bench workload (640x480, 5000
keypoints, LightGlue full depth).  Used for the per-launch time list and the `--set full` captures in this directory:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/capture_step.py
  ncu --set full --clock-control none --import-source on -k regex:k_flash_ws -c 1 -o gpurun_out/prof python profiles/capture_step.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gtsfm_b200 import synthetic as syn  # noqa: E402
from gtsfm_b200.pipeline import DeviceFrontEnd  # noqa: E402

fe = DeviceFrontEnd(syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "bench"), max_keypoints=5000)
frames, cal = syn.synthetic_sequence(3, 480, 640)
dev = [torch.from_numpy(f).cuda() for f in frames]
feats = [fe.detect(d) for d in dev]
for a, b in ((0, 1), (0, 2)):
    m, _ = fe.match(feats[a], feats[b])
    fe.verify(feats[a], feats[b], m, cal, cal, 4.0)
torch.cuda.synchronize()
print("launches", fe.ctx.launch_count())
