"""GPU: on-device batch sampling, the nonconstant-neighbourhood filter and classify_any
 against the numpy oracle -- and, where MuyGPyS is importable,
against the reference's own functions.  The random draws use torch's device generator, so the
sampled SETS are compared through their properties (sizes, uniqueness, class balance, filter
membership), everything deterministic bit for bit."""

import os
import sys

import numpy as np
import pytest
import torch

from oracle import numpy_oracle as O

from conftest import assert_close

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _blobs(seed, n=4000, d=5, classes=4, sep=1.6):
    rng = np.random.default_rng(seed)
    cent = rng.normal(0, sep, size=(classes, d))
    lab = rng.integers(0, classes, size=n)
    x = cent[lab] + rng.normal(size=(n, d))
    return x, lab


@pytest.mark.parametrize("host_api", [True, False], ids=["numpy-in", "device-in"])
def test_sample_batch_properties(host_api):
    from muygpys_b200.neighbors import NN_Wrapper
    from muygpys_b200.optimize.batch import sample_batch

    x, _ = _blobs(0)
    train = x if host_api else torch.as_tensor(x).cuda()
    nbrs = NN_Wrapper(train, 12)
    gen = torch.Generator(device="cuda").manual_seed(5)
    bi, bnn = sample_batch(nbrs, 500, len(x), generator=gen)
    if host_api:
        assert isinstance(bi, np.ndarray) and bi.dtype == np.int64
    else:
        assert bi.is_cuda and bnn.is_cuda
        bi, bnn = bi.cpu().numpy(), bnn.cpu().numpy()
    assert bi.shape == (500,) and bnn.shape == (500, 12)
    assert len(np.unique(bi)) == 500 and bi.min() >= 0 and bi.max() < len(x)
    want, _ = O.knn_batch(x, bi, 12)
    np.testing.assert_array_equal(bnn, want)
    # same seed, same draw; a different seed, a different one
    bi2, _ = sample_batch(nbrs, 500, len(x), generator=torch.Generator(device="cuda").manual_seed(5))
    bi3, _ = sample_batch(nbrs, 500, len(x), generator=torch.Generator(device="cuda").manual_seed(6))
    as_np = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy()  # noqa: E731
    np.testing.assert_array_equal(as_np(bi2), bi)
    assert not np.array_equal(as_np(bi3), bi)
    # batch_count >= train_count: everything, in order (batch.py:223-225)
    ball, _ = sample_batch(nbrs, 10 ** 6, len(x))
    np.testing.assert_array_equal(as_np(ball), np.arange(len(x)))


def test_full_filtered_and_balanced_batches():
    from muygpys_b200 import ops
    from muygpys_b200.neighbors import NN_Wrapper
    from muygpys_b200.optimize.batch import (full_filtered_batch, get_balanced_batch,
                                             sample_balanced_batch)

    x, lab = _blobs(1)
    k = 10
    nbrs = NN_Wrapper(x, k)
    want_idx, want_nn = O.full_filtered_batch(x, lab, k)
    got_idx, got_nn = full_filtered_batch(nbrs, lab)
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_nn, want_nn)
    assert 0 < len(want_idx) < len(x)
    # the kernel on its own, 1-D and one-hot labels
    nn_all, _ = O.knn_batch(x, np.arange(len(x)), k)
    onehot = -0.1 * np.ones((len(x), 4))
    onehot[np.arange(len(x)), lab] = 0.9
    for labels in (lab.astype(np.float64), onehot):
        mask = ops.nn_label_mask(torch.as_tensor(labels).cuda(), torch.as_tensor(nn_all).cuda())
        np.testing.assert_array_equal(mask.cpu().numpy(), O.nonconstant_mask(labels, nn_all))
    # balanced sample: per class min(available, batch_count / class_count), all from the
    # filtered set, classes in ascending order, no repeats
    batch_count = 400
    gen = torch.Generator(device="cuda").manual_seed(9)
    bi, bnn = sample_balanced_batch(nbrs, lab, batch_count, generator=gen)
    avail = {c: int(np.sum(lab[want_idx] == c)) for c in range(4)}
    counts = [min(avail[c], batch_count // 4) for c in range(4)]
    assert len(bi) == sum(counts) and len(np.unique(bi)) == len(bi)
    assert set(bi.tolist()) <= set(want_idx.tolist())
    np.testing.assert_array_equal(lab[bi], np.repeat(np.arange(4), counts))
    np.testing.assert_array_equal(bnn, nn_all[bi])
    # a class with fewer candidates than its share contributes all of them
    rare = np.where(lab == 3)[0][25:]
    lab2 = lab.copy()
    lab2[rare] = 2
    bi2, _ = sample_balanced_batch(nbrs, lab2, 2000)
    idx2, _ = O.full_filtered_batch(x, lab2, k)
    assert np.sum(lab2[bi2] == 3) == np.sum(lab2[idx2] == 3) <= 25
    # dispatcher (batch.py:58-64)
    a, _ = get_balanced_batch(nbrs, lab, len(lab) + 1)
    np.testing.assert_array_equal(a, want_idx)
    b, _ = get_balanced_batch(nbrs, lab, 40)
    assert len(b) <= 40


def test_classify_any_matches_oracle():
    from muygpys_b200.examples.classify import classify_any
    from muygpys_b200.gp import MuyGPS
    from muygpys_b200.gp.deformation import F2, Isotropy
    from muygpys_b200.gp.hyperparameter import Parameter
    from muygpys_b200.gp.kernels import RBF
    from muygpys_b200.gp.noise import HomoscedasticNoise
    from muygpys_b200.neighbors import NN_Wrapper

    x, lab = _blobs(2, n=3000, d=12, classes=3, sep=1.2)
    q, qlab = _blobs(2, n=300, d=12, classes=3, sep=1.2)
    onehot = -0.1 * np.ones((len(x), 3))
    onehot[np.arange(len(x)), lab] = 0.9
    k = 20
    model = MuyGPS(kernel=RBF(deformation=Isotropy(F2, Parameter(3.0))),
                   noise=HomoscedasticNoise(1e-3))
    pred, timing = classify_any(model, q, x, NN_Wrapper(x, k), onehot)
    want = O.classify_any(O.KERNEL_RBF, O.METRIC_F2, 3.0, 1e-3, x, onehot, q, k)
    assert set(timing) == {"nn", "agree", "pred"}
    assert_close(pred, want, 1e-10, "classify_any surrogate scores")
    mask = O.nonconstant_mask(onehot, O.knn_exact(x, q, k)[0])
    assert 0 < mask.sum() < len(q)
    np.testing.assert_array_equal(pred[~mask], want[~mask])  # agreed rows: exact label rows
    assert (pred.argmax(axis=1) == qlab).mean() > 0.8


def test_reference_batch_functions_agree():
    for extra in (os.path.join(ROOT, "oracle", "ref_shims"), os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(extra) and extra not in sys.path:
            sys.path.append(extra)
    pytest.importorskip("MuyGPyS", reason="reference not installed on this machine")
    from MuyGPyS.neighbors import NN_Wrapper as RefNN
    from MuyGPyS.optimize.batch import full_filtered_batch as ref_full

    from muygpys_b200.neighbors import NN_Wrapper
    from muygpys_b200.optimize.batch import full_filtered_batch

    x, lab = _blobs(3, n=2500, d=4)
    k = 8
    want_idx, want_nn = ref_full(RefNN(x, k, nn_method="exact", algorithm="ball_tree"), lab)
    got_idx, got_nn = full_filtered_batch(NN_Wrapper(x, k), lab)
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_nn, want_nn)
