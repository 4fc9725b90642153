"""`NN_Wrapper`: exact k-nearest-neighbour lookup on the GPU (S/neighbors.py:26-262).

Same surface as the reference (`get_nns`, `get_batch_nns`, attributes `train`,
`train_count`, `feature_count`, `nn_count`); results are int64 indices sorted by
ascending distance and SQUARED l2 distances.  The training set stays resident in
HBM; every query batch is one K2 launch.  Only `nn_method="exact"` with the
default Minkowski p=2 metric is built (HNSW is approximate and out of scope).
"""

from __future__ import annotations

from . import ops
from ._arrays import fdev, idev, like_input


class NN_Wrapper:
    def __init__(self, train, nn_count: int, nn_method: str = "exact", **kwargs):
        self._host_api = not hasattr(train, "is_cuda") or not train.is_cuda
        self._train_in = train
        t = fdev(train)
        if t.dim() == 1:
            t = t[:, None]
        self.train = t.contiguous()
        self.train_count, self.feature_count = self.train.shape
        self.nn_count = int(nn_count)
        self.nn_method = nn_method.lower()
        if self.nn_method != "exact":
            raise NotImplementedError(
                f"Nearest Neighbor algorithm {self.nn_method} is not implemented."
            )
        metric = kwargs.get("metric", "minkowski")
        p = kwargs.get("p", 2)
        if metric not in ("minkowski", "euclidean", "l2") or p != 2:
            raise NotImplementedError(
                f"only the l2 metric is built (got metric={metric!r}, p={p!r})"
            )
        # sklearn returns unsquared distances for metric="euclidean"; the
        # reference only squares for minkowski/p=2 (S/neighbors.py:246-250)
        self._squared = metric == "minkowski"
        # d <= 3: exact search on a uniform cell grid (bit-identical to brute force, orders of
        # magnitude fewer distance evaluations); otherwise brute force.  `algorithm="brute"`
        # (a sklearn keyword the reference forwards) forces the brute-force kernel.
        self._grid = None
        if (self.feature_count <= 3 and self.train_count >= 64
                and kwargs.get("algorithm", "auto") != "brute"):
            self._grid = ops.KnnGrid(self.train)

    def _query(self, samples, k: int):
        s = fdev(samples)
        if s.dim() == 1:
            s = s[:, None]
        if self._grid is not None:
            idx, d2 = self._grid.query(s, k)
        else:
            idx, d2 = ops.knn(self.train, s, k)
        if not self._squared:
            d2 = d2.sqrt()
        return idx, d2

    def get_nns(self, test):
        """(indices (q,k) int64, squared distances (q,k)) for arbitrary queries."""
        idx, d2 = self._query(test, self.nn_count)
        return like_input(idx, test), like_input(d2, test)

    def get_batch_nns(self, batch_indices):
        """Neighbours of training points: query k+1 and drop column 0, exactly as
        the reference does (S/neighbors.py:203-211) -- including its assumption that
        the self-match is the first column."""
        bi = idev(batch_indices)
        idx, d2 = self._query(self.train[bi], self.nn_count + 1)
        idx, d2 = idx[:, 1:].contiguous(), d2[:, 1:].contiguous()
        return like_input(idx, batch_indices), like_input(d2, batch_indices)

    # reference-internal name, kept because workflows call it
    def _get_nns(self, samples, nn_count: int):
        idx, d2 = self._query(samples, nn_count)
        return like_input(idx, samples), like_input(d2, samples)
