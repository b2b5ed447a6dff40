"""CPU: the retriever oracle against the golden made by the reference's own SimilarityRetriever."""
import numpy as np

from oracle import retriever_ref


def test_oracle_matches_reference_pairs(golden_dir):
    z = np.load(golden_dir / "retriever.npz")
    g = z["descriptors"]
    sim = retriever_ref.similarity_matrix(g)
    ref = z["sim"]
    assert np.abs(sim - ref).max() < 2e-6
    assert np.all(np.tril(ref, -50) == 0)  # the reference fills the upper block triangle only
    for c in range(4):
        k, ms = z[f"case_{c}"]
        pairs = retriever_ref.similarity_pairs(ref.copy(), int(k), float(ms))
        assert pairs == [tuple(p) for p in z[f"pairs_{c}"].tolist()], c


def test_oracle_edge_cases():
    sim = np.array([[1.0, 0.5, 0.5], [0.5, 1.0, 0.05], [0.5, 0.05, 1.0]], np.float32)
    assert retriever_ref.similarity_pairs(sim, 1, 0.1) == [(0, 1)]  # tie -> lower index; (1,2) below min_score
    assert retriever_ref.similarity_pairs(sim, 5, 0.0) == [(0, 1), (0, 2), (1, 2)]
    assert retriever_ref.similarity_pairs(sim[:1, :1], 3, 0.1) == []
