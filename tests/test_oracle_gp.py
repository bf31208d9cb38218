"""CPU tests of the GP oracle (oracle/gp_oracle.py): the restatement of GPy's exact GP that every GPU parity test
checks against.  The reference pins nothing for this path ("parity unpinned"), so the oracle is pinned
here from three independent sides:
  1. scikit-learn (named beside GPy at "Kernel Selection/README.md":9) for every family both libraries share,
     at fixed hyper-parameters: LML, gradient, predictive mean and variance (SURVEY.md App. A.6 recipe);
  2. central finite differences of the LML for every family and for + / * composites;
  3. the survey's anchors on the reference's only data fixture, core_navigation/script/slipVal.csv.
"""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process import kernels as sk

from oracle import gp_oracle as go

NOISE = 2e-3


def series(n=60, seed=0):
    rng = np.random.default_rng(seed)
    x = 20.0 + np.arange(n) + rng.uniform(-0.2, 0.2, n)
    y = 0.02 + 0.05 * np.sin(2 * np.pi * np.arange(n) / 37.0) + 0.03 * rng.standard_normal(n)
    return x, y


# (oracle expression, theta, scikit-learn twin).  GPy -> sklearn parameter maps: SURVEY.md App. A.3.
TWINS = [
    ("rbf", [0.7, 9.0], lambda: sk.ConstantKernel(0.7) * sk.RBF(9.0)),
    ("mat32", [0.4, 6.0], lambda: sk.ConstantKernel(0.4) * sk.Matern(6.0, nu=1.5)),
    ("mat52", [0.4, 6.0], lambda: sk.ConstantKernel(0.4) * sk.Matern(6.0, nu=2.5)),
    ("ratquad", [0.3, 8.0, 1.7], lambda: sk.ConstantKernel(0.3) * sk.RationalQuadratic(8.0 / np.sqrt(1.7), 1.7)),
    ("stdperiodic", [0.5, 37.0, 1.2], lambda: sk.ConstantKernel(0.5) * sk.ExpSineSquared(2 * 1.2, 37.0)),
    ("linear", [1e-4], lambda: sk.ConstantKernel(1e-4) * sk.DotProduct(0.0)),
    ("bias", [0.2], lambda: sk.ConstantKernel(0.2)),
    ("rbf+stdperiodic", [0.01, 10.0, 0.0025, 37.0, 1.0],
     lambda: sk.ConstantKernel(0.01) * sk.RBF(10.0) + sk.ConstantKernel(0.0025) * sk.ExpSineSquared(2.0, 37.0)),
    ("rbf*linear", [0.5, 12.0, 1e-3],
     lambda: (sk.ConstantKernel(0.5) * sk.RBF(12.0)) * (sk.ConstantKernel(1e-3) * sk.DotProduct(0.0))),
    ("mat32+rbf*stdperiodic", [0.02, 5.0, 0.3, 30.0, 0.4, 37.0, 0.9],
     lambda: sk.ConstantKernel(0.02) * sk.Matern(5.0, nu=1.5) +
     (sk.ConstantKernel(0.3) * sk.RBF(30.0)) * (sk.ConstantKernel(0.4) * sk.ExpSineSquared(1.8, 37.0))),
]


@pytest.mark.parametrize("expr,theta,twin", TWINS, ids=[t[0] for t in TWINS])
def test_oracle_matches_sklearn(expr, theta, twin):
    x, y = series()
    xs = x[-1] + 1.0 + np.arange(40.0)
    e = go.KernelExpr(expr)
    inf = go.inference(e, theta, NOISE, x, y)
    mu, var = go.predict(e, theta, NOISE, x, y, xs, inf)
    gpr = GaussianProcessRegressor(kernel=twin() + sk.WhiteKernel(NOISE), optimizer=None, alpha=go.JITTER)
    gpr.fit(x[:, None], y)
    mu_s, sd_s = gpr.predict(xs[:, None], return_std=True)
    assert abs(inf.lml - gpr.log_marginal_likelihood_value_) < 1e-9 * max(1.0, abs(inf.lml))
    assert np.max(np.abs(mu - mu_s)) < 1e-9
    assert np.max(np.abs(var - sd_s ** 2)) < 1e-9


def test_oracle_gradient_matches_sklearn_rbf():
    """sklearn differentiates with respect to log(theta): d/dlog(t) = t d/dt."""
    x, y = series(50, 2)
    th = [0.7, 9.0]
    inf = go.inference(go.KernelExpr("rbf"), th, NOISE, x, y, want_grad=True)
    kern = sk.ConstantKernel(th[0]) * sk.RBF(th[1]) + sk.WhiteKernel(NOISE)
    gpr = GaussianProcessRegressor(kernel=kern, optimizer=None, alpha=go.JITTER).fit(x[:, None], y)
    lml, g = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    ours = inf.grad * np.array(th + [NOISE])
    assert np.max(np.abs(ours - g) / np.maximum(1.0, np.abs(g))) < 1e-7


FD_CASES = [
    ("rbf", [0.7, 9.0]), ("mat32", [0.4, 6.0]), ("mat52", [0.4, 6.0]), ("ratquad", [0.3, 8.0, 1.7]),
    ("stdperiodic", [0.5, 37.0, 1.2]), ("brownian", [0.01]), ("linear", [1e-4]), ("bias", [0.2]), ("white", [0.01]),
    ("rbf*brownian", [1.0, 1.0, 1.0]), ("rbf+stdperiodic", [0.01, 10.0, 0.0025, 37.0, 1.0]),
    ("(rbf+linear)*brownian+white", [0.02, 12.0, 1e-6, 0.03, 2e-4]),
    ("ratquad+stdperiodic*rbf", [0.01, 8.0, 1.5, 0.5, 37.0, 1.2, 0.01, 50.0]),
]


@pytest.mark.parametrize("expr,theta", FD_CASES, ids=[c[0] for c in FD_CASES])
def test_oracle_gradient_finite_difference(expr, theta):
    x, y = series(40, 3)
    e = go.KernelExpr(expr)
    p = np.array(theta + [5e-3])
    g = go.inference(e, p[:-1], p[-1], x, y, want_grad=True).grad
    for j in range(p.size):
        h = 1e-6 * max(1e-3, abs(p[j]))
        pp, pm = p.copy(), p.copy()
        pp[j] += h
        pm[j] -= h
        fd = (go.inference(e, pp[:-1], pp[-1], x, y).lml - go.inference(e, pm[:-1], pm[-1], x, y).lml) / (2 * h)
        assert abs(fd - g[j]) < 2e-5 * max(1.0, abs(g[j]), abs(fd)), (expr, j, fd, g[j])


def test_slipval_anchors(slipval):
    """SURVEY.md section 8c anchors on the reference's own fixture (train = first int(0.9*199) = 179 rows, all
    hyper-parameters and the noise at GPy's initial value 1.0)."""
    t, s = slipval
    assert t.size == 199
    xtr, ytr = go.split_train(t, s)
    assert xtr.size == 179
    lml_dep = go.inference(go.KernelExpr("rbf*brownian"), [1.0, 1.0, 1.0], 1.0, xtr, ytr).lml
    lml_rbf = go.inference(go.KernelExpr("rbf"), [1.0, 1.0], 1.0, xtr, ytr).lml
    assert abs(lml_dep - (-223.18875545549918)) < 1e-9
    assert abs(lml_rbf - (-182.7789464331909)) < 1e-9
    gpr = GaussianProcessRegressor(kernel=sk.ConstantKernel(1.0) * sk.RBF(1.0) + sk.WhiteKernel(1.0), optimizer=None,
                                   alpha=go.JITTER).fit(xtr[:, None], ytr)
    assert abs(lml_rbf - gpr.log_marginal_likelihood_value_) < 1e-10


def test_prediction_grid_and_callback_shapes(slipval):
    """gp_slip_node.py:45,57-61: grid = arange(min, max+600, 1); outputs drop the first len(X) entries; sigma = 2 sd."""
    t, s = slipval
    grid = go.prediction_grid(t)
    assert grid[0] == t.min() and grid.size == int(np.ceil(t.max() + 600 - t.min()))
    e = go.KernelExpr("rbf*brownian")
    mean, sigma = go.gp_slip_callback(t, s, e, theta=[1.0, 1.0, 1.0], noise=1.0)
    assert mean.size == grid.size - t.size == sigma.size
    xtr, ytr = go.split_train(t, s)
    mu, var = go.predict(e, [1.0, 1.0, 1.0], 1.0, xtr, ytr, grid)
    assert np.array_equal(mean, mu[t.size:]) and np.array_equal(sigma, 2.0 * np.sqrt(var[t.size:]))
    # far from the data an RBF*Brownian prior reverts to mean 0 and the band keeps widening (Kernel Selection/docs/ours.jpg)
    assert abs(mean[-1]) < 1e-6 and sigma[-1] > sigma[0]


def test_pointwise_predict_equals_block_predict():
    """The reference calls m.predict once per point (gp_slip_node.py:47-50); one M-column dtrtrs gives the same numbers."""
    x, y = series(30, 4)
    xs = x[-1] + 1.0 + np.arange(25.0)
    e = go.KernelExpr("rbf*brownian")
    a = go.predict(e, [0.5, 8.0, 0.05], 1e-3, x, y, xs)
    b = go.predict_pointwise(e, [0.5, 8.0, 0.05], 1e-3, x, y, xs)
    assert np.allclose(a[0], b[0], rtol=0, atol=1e-14) and np.allclose(a[1], b[1], rtol=0, atol=1e-14)


def test_variance_floor_and_noise():
    """posterior.py clips the latent variance at 1e-15 BEFORE the noise is added."""
    x = np.arange(5.0)
    y = np.zeros(5)
    e = go.KernelExpr("bias")
    mu, var = go.predict(e, [1.0], 0.0, x, y, x)          # Ky = 11' + 1e-8 I: latent variance ~1e-8/5 > floor
    assert np.all(var > 0)
    mu, var = go.predict(e, [1e-20], 0.25, x, y, x)
    assert np.allclose(var, 0.25 + 1e-15, rtol=1e-12, atol=0) or np.all(var >= 0.25)


def test_jitchol_ladder():
    A = np.ones((6, 6))                       # rank one: dpotrf fails, first jitter = mean(diag)*1e-6 succeeds
    L, jit = go.jitchol(A)
    assert jit == pytest.approx(1e-6) and np.allclose(L @ L.T, A + jit * np.eye(6))
    with pytest.raises(go.NotPositiveDefinite):
        go.jitchol(-np.eye(3))


def test_optimizer_improves_lml(slipval):
    t, s = slipval
    xtr, ytr = go.split_train(t, s)
    e = go.KernelExpr("rbf*brownian")
    lml0 = go.inference(e, [1.0, 1.0, 1.0], 1.0, xtr, ytr).lml
    th, noise, lml, nev = go.optimize(e, xtr, ytr, max_iters=200)
    assert lml > lml0 + 100 and np.all(th > 0) and noise > 0 and nev <= 200
    assert abs(go.inference(e, th, noise, xtr, ytr).lml - lml) < 1e-8 * abs(lml)


def test_softplus_transform_roundtrip():
    t = np.array([1e-6, 0.3, 1.0, 20.0, 50.0])
    assert np.allclose(go.softplus(go.softplus_inv(t)), t, rtol=1e-12)
    z = go.softplus_inv(t)
    h = 1e-6
    fd = (go.softplus(z + h) - go.softplus(z - h)) / (2 * h)
    assert np.allclose(fd, go.softplus_gradfactor(t), rtol=1e-5)
