import sys; sys.path.insert(0, ".")
import time, numpy as np, torch
from gtsfm_b200 import synthetic as syn
from gtsfm_b200.detector_descriptor import B200SuperPointDetectorDescriptor
from gtsfm_b200.gtsfm_api import Cal3Bundler, Image
from gtsfm_b200.matcher import B200LightGlueMatcher
from gtsfm_b200.verifier import B200Ransac
frames, cal = syn.synthetic_sequence(24, 480, 640)
sp_sd, lg_sd = syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "bench")
det = B200SuperPointDetectorDescriptor(max_keypoints=5000, weights_path=sp_sd)
mat = B200LightGlueMatcher("superpoint", weights_path=lg_sd)
ver = B200Ransac(True, 4.0)
calib = Cal3Bundler(cal[0], 0, 0, cal[1], cal[2])
feats = [det.detect_and_describe(Image(f)) for f in frames[:6]]
T = {"match": 0.0, "verify": 0.0, "detect": 0.0}
n = 0
for rep in range(2):
    for i in range(5):
        (k0, d0), (k1, d1) = feats[i], feats[i + 1]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m = mat.match(k0, k1, d0, d1, (480, 640, 3), (480, 640, 3))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ver.verify(k0, k1, m, calib, calib)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if rep:
            T["match"] += t1 - t0; T["verify"] += t2 - t1; n += 1
t0 = time.perf_counter()
for f in frames[6:16]:
    det.detect_and_describe(Image(f))
T["detect"] = (time.perf_counter() - t0) / 10
print({k: (v / n if k != "detect" else v) * 1e3 for k, v in T.items()}, "ms per call; matches", len(m))
# inside match: host->device copy cost alone
x = np.random.rand(5000, 256).astype(np.float32)
d = torch.empty((5000, 256), device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    d.copy_(torch.from_numpy(x)); torch.cuda.synchronize()
print("pageable 5 MB H2D:", (time.perf_counter() - t0) / 20 * 1e3, "ms")
eng = mat._engine
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    (k0, d0), (k1, d1) = feats[i], feats[i + 1]
    mat.match(k0, k1, d0, d1, (480, 640, 3), (480, 640, 3))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
