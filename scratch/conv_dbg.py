import sys; sys.path.insert(0, ".")
import os
os.environ["B2_CONV_DBG"] = "1"
import numpy as np, torch
from gtsfm_b200 import synthetic as syn
from gtsfm_b200.pipeline import DeviceFrontEnd
fe = DeviceFrontEnd(syn.superpoint_state_dict(0), syn.lightglue_state_dict(2), max_keypoints=5000)
rng = np.random.default_rng(0)
img = torch.from_numpy(syn.synthetic_frame(0, 480, 640) if hasattr(syn, "synthetic_frame") else rng.integers(0, 255, (480, 640), dtype=np.uint8)).cuda()
for _ in range(3):
    fe.detect(img)
torch.cuda.synchronize()
d = fe.ctx.debug_fetch("conv_dbg", 12 * 148 * 8).reshape(12, 148, 8)
names = {1: "1b", 2: "2a", 3: "2b", 4: "3a", 5: "3b", 6: "4a", 7: "4b", 8: "Pa", 10: "Da"}
for li, nm in names.items():
    x = d[li]
    act = x[:, 7] > 0
    x = x[act]
    def dl(a, b):
        v = (x[:, a] - x[:, b]) % (1 << 24)
        return v
    t0 = x[:, 0].min()
    print(f"{nm}: ctas {act.sum()} tiles/cta {x[:,7].min():.0f}-{x[:,7].max():.0f} | start spread {(x[:,0]-t0).max()/1e3:.2f} us | setup {np.median(dl(1,0))/1e3:.2f} | first-mma after setup {np.median(dl(2,1))/1e3:.2f} (max {dl(2,1).max()/1e3:.2f}) | mma span {np.median(dl(3,2))/1e3:.2f} (max {dl(3,2).max()/1e3:.2f}) | first acc ready after first mma {np.median(dl(4,2))/1e3:.2f} | last acc -> end {np.median(dl(6,5))/1e3:.2f} | total {np.median(dl(6,0))/1e3:.2f} (max {dl(6,0).max()/1e3:.2f}) | kernel {((x[:,6]-t0)%(1<<24)).max()/1e3:.2f}")
