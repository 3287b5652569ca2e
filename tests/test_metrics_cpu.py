"""The Chamfer oracle against an independent float64 brute force (well-separated random clouds:
the fp32 rounding of the kernel's distance cannot change a nearest neighbour) and its tie rule."""
import numpy as np

from oracle import metrics as OM


def test_chamfer_oracle_matches_float64_brute_force():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((2, 300, 3)).astype(np.float32)
    b = rng.standard_normal((2, 257, 3)).astype(np.float32)
    d1, d2, i1, i2 = OM.chamfer_forward(a, b)
    for k in range(2):
        D = ((a[k].astype(np.float64)[:, None] - b[k].astype(np.float64)[None]) ** 2).sum(-1)
        assert np.array_equal(i1[k], D.argmin(1)) and np.array_equal(i2[k], D.argmin(0))
        assert np.allclose(d1[k], D.min(1), rtol=1e-5) and np.allclose(d2[k], D.min(0), rtol=1e-5)
    cd = OM.pairwise_cd(a, b)
    assert cd.shape == (2, 2)
    assert abs(cd[1, 0] - (OM._nn(a[1], b[0])[0].mean() + OM._nn(b[0], a[1])[0].mean())) < 1e-6


def test_chamfer_oracle_tie_rule_lowest_index():
    q = np.zeros((1, 4, 3), np.float32)
    c = np.zeros((1, 6, 3), np.float32)
    c[0, :, 0] = [2, 1, 1, 3, 1, 0.5]        # candidates 1, 2, 4 tie at distance 1 until index 5 wins
    d1, _, i1, _ = OM.chamfer_forward(q, c)
    assert (i1 == 5).all() and np.allclose(d1, 0.25)
    c[0, 5, 0] = 1
    _, _, i1, _ = OM.chamfer_forward(q, c)
    assert (i1 == 1).all()


def test_emd_oracle_properties():
    """Identical clouds cost ~0; for a rigid shift t the optimal transport costs |t|^2 per point, the
    approximate scheme (soft assignment annealed over ten levels) can only do worse, and the cost
    grows with |t|."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal((1, 128, 3)).astype(np.float32) * 0.3
    assert OM.emd_approx(a, a)[0] / 128 < 1e-6
    prev = 0.0
    for t in (0.1, 0.5, 1.0):
        c = OM.emd_approx(a, a + np.array([t, 0, 0], np.float32))[0] / 128
        assert c >= t * t * 0.999 and c < 6 * t * t and c > prev
        prev = c
