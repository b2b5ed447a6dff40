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


def test_mat_checkpoint_parser(tmp_path):
    """weights.load_netvlad_mat against a MATLAB file laid out like the reference's checkpoint (thirdparty/hloc/netvlad.py:115-160):
    net.layers[i].weights = [S x S x IN x OUT, OUT] at the backbone's Conv2d positions, layers[30] = NetVLAD (D x K scores, D x K
    negated centres), layers[33] = whitening (1 x 1 x IN x OUT, OUT), net.meta.normalization.averageImage."""
    import scipy.io

    rng = np.random.default_rng(0)
    dims = {i: (ci if ci <= 3 else 4, 5) for i, ci in zip(weights.NETVLAD_CONV_IDX, [3] + [4] * 12)}
    layers = []
    for i in range(34):
        if i in dims:
            ci, co = dims[i]
            layers.append({"type": "conv", "weights": np.array([rng.standard_normal((3, 3, ci, co)), rng.standard_normal(co)], dtype=object)})
        elif i == 30:
            layers.append({"type": "vlad", "weights": np.array([rng.standard_normal((6, 7)), rng.standard_normal((6, 7))], dtype=object)})
        elif i == 33:
            layers.append({"type": "whiten", "weights": np.array([rng.standard_normal((1, 1, 42, 9)), rng.standard_normal((9, 1))], dtype=object)})
        else:
            layers.append({"type": "relu", "weights": np.zeros(0)})
    mean = np.zeros((2, 2, 3))
    mean[0, 0] = [123.0, 117.0, 104.0]
    path = tmp_path / "fake_struct.mat"
    scipy.io.savemat(str(path), {"net": {"layers": np.array(layers, dtype=object), "meta": {"normalization": {"averageImage": mean}}}})
    sd = weights.load_netvlad_mat(path)
    mat = scipy.io.loadmat(str(path), struct_as_record=False, squeeze_me=True)
    for i in weights.NETVLAD_CONV_IDX:
        w = np.asarray(mat["net"].layers[i].weights[0], np.float64)
        assert sd[f"backbone.{i}.weight"].shape == (w.shape[3], w.shape[2], 3, 3)
        assert np.allclose(sd[f"backbone.{i}.weight"][1, 2, 0, 1], w[0, 1, 2, 1]) and np.allclose(sd[f"backbone.{i}.bias"], np.asarray(mat["net"].layers[i].weights[1], np.float64))
    assert sd["netvlad.score_proj.weight"].shape == (7, 6, 1) and np.allclose(sd["netvlad.score_proj.weight"][2, 3, 0], np.asarray(mat["net"].layers[30].weights[0], np.float64)[3, 2])
    assert np.allclose(sd["netvlad.centers"], -np.asarray(mat["net"].layers[30].weights[1], np.float64))
    assert sd["whiten.weight"].shape == (9, 42) and np.allclose(sd["whiten.weight"][4, 5], np.asarray(mat["net"].layers[33].weights[0], np.float64)[5, 4])
    assert sd["whiten.bias"].shape == (9,) and sd["mean"].tolist() == [123.0, 117.0, 104.0]
    assert set(sd) == set(weights.NETVLAD_ORDER)
