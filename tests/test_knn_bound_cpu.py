"""CPU check of the rounding bound behind the KNN pre-filter's certificate (csrc/knn_gram.cu).

The filter ranks by S~ = |q|^2 + |x|^2 - 2 q.x and the refine step certifies a result when the
exact k-th distance is below (tau - 2 gamma (|q|^2 + max|x|^2)) (1 - gamma) with
gamma = 8 (d + 8) 2^-53.  Here the same quantities are formed in float64 on the CPU (dot products
by BLAS, norms by pairwise summation -- different summation orders from the device's, same error
model) and compared with the brute-force arithmetic: the observed |S~ - S| must stay well inside
the margin the kernel allows, also for data far from the origin where the identity cancels."""

import numpy as np
import pytest


def brute_force_sq(q, x):
    s = np.zeros((q.shape[0], x.shape[0]))
    for f in range(x.shape[1]):  # separately rounded subtract / multiply / add, feature order
        s += (q[:, f:f + 1] - x[None, :, f]) ** 2
    return s


@pytest.mark.parametrize("d,offset", [(4, 0.0), (33, 0.0), (784, 0.0), (33, 1e4), (784, 1e3)])
def test_gram_identity_error_is_inside_the_certificate_margin(d, offset):
    rng = np.random.default_rng(d + int(offset))
    x = rng.normal(size=(3000, d)) + offset
    q = rng.normal(size=(40, d)) + offset
    qn = np.einsum("ij,ij->i", q, q)
    xn = np.einsum("ij,ij->i", x, x)
    s_tilde = qn[:, None] + xn[None, :] - 2.0 * (q @ x.T)
    s_exact = brute_force_sq(q, x)
    gamma = 8.0 * (d + 8) * 2.0 ** -53
    margin = 2.0 * gamma * (qn[:, None] + xn.max()) + gamma * np.abs(s_tilde)
    err = np.abs(s_tilde - s_exact)
    assert (err <= margin).all()
    # the margin is conservative (>= 4x the worst observed error) yet tiny next to the distances
    assert err.max() <= 0.25 * margin.min()
    if offset == 0.0:
        assert margin.max() < 1e-9 * np.median(s_exact)


@pytest.mark.parametrize("d", [12, 784])
def test_centred_gram_distances_keep_eleven_digits_above_the_fixup_ratio(d):
    """The d > 8 assembly of the fused kernel (csrc/gram.cuh) takes |u_i - u_j|^2 from Gram
    tiles of the query-centred rows and recomputes by direct differences every pair with
    |u_i - u_j|^2 < (|u_i|^2 + |u_j|^2) / 256.  Above that ratio the identity must be good to
    ~1e-11 relative (the fused outputs are compared at 1e-10): checked here in float64 on
    neighbourhoods far from the origin, where the uncentred identity would lose everything."""
    rng = np.random.default_rng(d)
    k = 40
    centre = rng.normal(size=d) * 1e3
    x = centre + rng.normal(size=(k, d))          # a neighbourhood of radius ~sqrt(d)
    x[1] = x[0] + 0.05 * rng.normal(size=d)       # a close pair, still above the ratio
    q = centre + 0.3 * rng.normal(size=d)
    u = x - q
    g = u @ u.T
    n = np.diag(g)
    s_gram = n[:, None] + n[None, :] - 2.0 * g
    s_exact = brute_force_sq(x, x)
    iu = np.triu_indices(k, 1)
    keep = s_gram[iu] >= (n[:, None] + n[None, :])[iu] / 256.0
    assert keep.sum() > 0.9 * keep.size
    rel = np.abs(s_gram[iu] - s_exact[iu])[keep] / s_exact[iu][keep]
    assert rel.max() < 2e-11
    # without centring the same identity is useless at this offset
    g0 = x @ x.T
    n0 = np.diag(g0)
    rel0 = np.abs((n0[:, None] + n0[None, :] - 2.0 * g0)[iu] - s_exact[iu]) / s_exact[iu]
    assert rel0.max() > 1e-9
