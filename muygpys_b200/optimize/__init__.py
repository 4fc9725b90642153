"""`muygpys_b200.optimize`: the fused objective factory.  The outer optimisers are the
reference's own (`MuyGPyS.optimize.L_BFGS_B_optimize` / `Bayes_optimize`): hand them to
`muygpys_b200.examples.from_indices.optimize_from_indices(opt_fn=...)`."""

from ..objective import make_fused_loo_crossval_fn, make_loo_crossval_fn  # noqa: F401
