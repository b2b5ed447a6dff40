"""CPU: the NetVLAD oracle against the golden made by the reference module's own forward (oracle/make_golden.py golden_netvlad)."""
import numpy as np

from gtsfm_b200 import synthetic as syn, weights
from oracle import netvlad_ref


def test_oracle_equals_reference_forward(golden_dir):
    z = np.load(golden_dir / "netvlad.npz")
    sd = syn.netvlad_state_dict(3)
    shapes = [tuple(z[f"shape_{i}"]) for i in range(3)]
    imgs = [np.ascontiguousarray(syn.synthetic_frame(40 + i, h, w).transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
            for i, (h, w) in enumerate(shapes)]
    for i, im in enumerate(imgs):
        d = netvlad_ref.netvlad_forward(sd, im[None])[0]
        assert np.abs(d - z[f"desc_{i}"]).max() < 1e-6
    both = netvlad_ref.netvlad_forward(sd, np.stack([imgs[0], imgs[2]]))
    assert np.abs(both - z["desc_batch_0_2"]).max() < 1e-6
    blob = weights.pack_netvlad(sd)  # the blob layout the library expects
    assert blob.dtype == np.float32 and blob[-3:].tolist() == sd["mean"].tolist()
