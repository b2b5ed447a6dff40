"""ncu workload for the retrieval front: NetVLAD on two 640x480 images (one batch) + similarity / top-k over
2000 descriptors of 4096 floats.   ncu --clock-control none --metrics gpu__time_duration.sum --csv python profiles/capture_r02_retrieval.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gtsfm_b200 import synthetic as syn  # noqa: E402
from gtsfm_b200.global_descriptor import NetVLADEngine  # noqa: E402
from gtsfm_b200.retriever import B200SimilarityRetriever  # noqa: E402

eng = NetVLADEngine(syn.netvlad_state_dict(3))
x = torch.stack([torch.from_numpy(np.ascontiguousarray(syn.synthetic_frame(40 + i, 480, 640).transpose(2, 0, 1))).float() / 255 for i in range(2)]).cuda()
d = eng.describe_dev(x)
torch.cuda.synchronize()
rng = np.random.default_rng(0)
g = rng.standard_normal((2000, 4096)).astype(np.float32)
g /= np.linalg.norm(g, axis=1, keepdims=True)
pairs = B200SimilarityRetriever(num_matched=10, min_score=0.0).get_image_pairs(list(g), [""] * len(g))
print("netvlad", tuple(d.shape), float(d[0] @ d[1]), "pairs", len(pairs))
