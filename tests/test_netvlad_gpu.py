"""GPU: NetVLAD global descriptor (b2_netvlad_*) against the golden made by the reference module's own forward and the oracle.
Tolerance: 2e-5 absolute on unit-norm 4096-vectors (elements ~1.5e-2; two different images differ by ~3e-3 per element)."""
import numpy as np
import pytest

from gtsfm_b200 import synthetic as syn
from oracle import netvlad_ref

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _images(shapes):
    return [np.ascontiguousarray(syn.synthetic_frame(40 + i, h, w).transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
            for i, (h, w) in enumerate(shapes)]


@pytest.fixture(scope="module")
def engine(b200_ctx):
    from gtsfm_b200.global_descriptor import NetVLADEngine

    return NetVLADEngine(syn.netvlad_state_dict(3), ctx=b200_ctx)


def test_golden_descriptors(engine, golden_dir):
    z = np.load(golden_dir / "netvlad.npz")
    imgs = _images(((96, 128), (120, 168), (96, 128)))
    for i, im in enumerate(imgs):
        assert tuple(z[f"shape_{i}"]) == im.shape[1:]
        d = engine.describe(im[None])[0]
        assert abs(np.linalg.norm(d) - 1) < 1e-5
        assert np.abs(d - z[f"desc_{i}"]).max() < TOL, (i, np.abs(d - z[f"desc_{i}"]).max())
    both = engine.describe(np.stack([imgs[0], imgs[2]]))  # a batch shares the whitening GEMM
    assert np.abs(both - z["desc_batch_0_2"]).max() < TOL


def test_other_sizes_against_oracle(engine):
    sd = syn.netvlad_state_dict(3)
    for im in _images(((64, 80), (200, 136))):
        want = netvlad_ref.netvlad_forward(sd, im[None])[0]
        got = engine.describe(im[None])[0]
        assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


def test_plugin_contract_and_retrieval(engine, golden_dir, tmp_path):
    import torch

    from gtsfm_b200.global_descriptor import B200NetVLADGlobalDescriptor
    from gtsfm_b200.retriever import B200SimilarityRetriever

    with pytest.raises(FileNotFoundError):
        B200NetVLADGlobalDescriptor(weights_path=tmp_path / "missing.mat")
    g = B200NetVLADGlobalDescriptor(weights_path=syn.netvlad_state_dict(3))
    g._engine = engine  # reuse the loaded weights (the whitening layer is 537 MB)
    resize, batch = g.get_preprocessing_transforms()
    frames = [syn.synthetic_frame(40 + i, 96, 128) for i in (0, 2, 0)]
    x = batch(torch.stack([resize(f) for f in frames]))
    descs = g.describe_batch(x)
    z = np.load(golden_dir / "netvlad.npz")
    assert len(descs) == 3 and descs[0].shape == (4096,)
    assert np.abs(descs[0] - z["desc_0"]).max() < TOL and np.abs(descs[1] - z["desc_2"]).max() < TOL
    assert np.abs(descs[0] - descs[2]).max() < 1e-6
    pairs = B200SimilarityRetriever(num_matched=1, min_score=0.3).get_image_pairs(descs, ["a", "b", "c"])
    assert pairs == [(0, 2), (1, 2)]  # frame 0 twice: its copy is its best partner
