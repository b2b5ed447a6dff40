import sys; sys.path.insert(0, ".")
import time, numpy as np, torch
from gtsfm_b200 import synthetic as syn
from gtsfm_b200.global_descriptor import NetVLADEngine
sd = syn.netvlad_state_dict(3)
eng = NetVLADEngine(sd)
for (h, w, b) in [(480, 640, 8), (760, 1013, 4)]:
    x = torch.rand((b, 3, h, w), device="cuda")
    eng.describe_dev(x); torch.cuda.synchronize()
    t = time.perf_counter(); eng.describe_dev(x); torch.cuda.synchronize(); dt = time.perf_counter() - t
    flop = 0.0
    hh, ww = h, w
    for (idx, ci, co) in syn.NETVLAD_CONVS:
        flop += 2.0 * 9 * hh * ww * ci * co
        if idx in syn.NETVLAD_POOL_AFTER: hh //= 2; ww //= 2
    print(f"{h}x{w} batch {b}: {dt / b * 1e3:.2f} ms per image, {b / dt:.1f} img/s, backbone {flop * b / dt / 1e12:.1f} TFLOP/s algorithmic")
from oracle import netvlad_ref
torch.set_num_threads(32)
im = np.random.rand(1, 3, 480, 640).astype(np.float32)
t = time.perf_counter(); netvlad_ref.netvlad_forward(sd, im); print("oracle (torch CPU, 32 threads) 480x640:", time.perf_counter() - t, "s")
