import sys; sys.path.insert(0, ".")
import time, numpy as np, torch
from gtsfm_b200.retriever import B200SimilarityRetriever
for n, dim in [(1000, 4096), (5000, 4096), (10000, 4096), (10000, 8448)]:
    rng = np.random.default_rng(0)
    g = rng.standard_normal((n, dim)).astype(np.float32); g /= np.linalg.norm(g, axis=1, keepdims=True)
    r = B200SimilarityRetriever(20, 0.0)
    r.similarity_and_partners(g, want_sim=False)
    t = time.perf_counter(); r.similarity_and_partners(g, want_sim=False); dt = time.perf_counter() - t
    t = time.perf_counter(); gt = torch.from_numpy(g); s = gt @ gt.T; torch.topk(s, 20, dim=1); dc = time.perf_counter() - t
    print(f"n={n} dim={dim}: b200 {dt*1e3:.1f} ms (host buffers in, partners out), torch cpu {dc*1e3:.1f} ms", flush=True)
