import sys; sys.path.insert(0,'.')
import numpy as np
from gtsfm_b200 import synthetic as syn, _lib
from gtsfm_b200.detector_descriptor import SuperPointEngine
from gtsfm_b200.matcher import LightGlueEngine
ctx=_lib.Context(0)
sp = SuperPointEngine(syn.superpoint_state_dict(0), ctx=ctx)
lg = LightGlueEngine(syn.lightglue_state_dict(2, "sharp"), ctx=ctx)
def feats(gray, k=5000):
    xy, sc = sp.detect(gray)
    sel = np.argpartition(-sc, k)[:k] if len(xy) > k else np.arange(len(xy))
    return xy[sel], sc[sel], sp.describe(xy[sel])
g1 = np.load("tests/golden/superpoint_lund1.npz")["gray"]; g2 = np.load("tests/golden/superpoint_lund2.npz")["gray"]
for name,(A,B),fxn in [("lund12",(g1,g2),"lightglue_lund_1_2"),("crops",(np.ascontiguousarray(g1[0:1000,0:700]),np.ascontiguousarray(g1[40:1040,24:724])),"pipeline_lund_crops_sharp")]:
    fa, fb = feats(A), feats(B)
    fx = np.load(f"tests/golden/{fxn}.npz")
    m, sc = lg.match(fa[0], fa[2], fb[0], fb[2], return_scores=True)
    ref = fx["matches"]
    sm = set(map(tuple,m.tolist())); sr=set(map(tuple,ref.tolist()))
    print(name, "gpu", len(m), "ref", len(ref), "common", len(sm&sr), "only gpu", sorted(sm-sr), "only ref", sorted(sr-sm))
    d = {tuple(r):s for r,s in zip(m.tolist(), sc.tolist())}
    print("  scores of gpu-only:", [d[k] for k in sorted(sm-sr)])
