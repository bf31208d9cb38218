"""CPU oracle for the GP slip-regression half of the hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  The product path
(corenav_gp_b200) never imports anything from oracle/.

PARITY UNPINNED: the arithmetic of this path lives in GPy (PyPI "GPy", SheffieldML; the reference
names no version - `import GPy` at core_navigation/script/gp_slip_node.py:3 is the only pin; the
.pyc artefacts next to it date it to GPy 1.9.x), which is neither vendored in /root/reference nor
installable here, and the reference holds no golden vector or test for this path
(gp_predictor/CMakeLists.txt:201-211 is a commented-out template).  This file therefore restates
GPy's published exact-GP algorithm and is cross-validated against scikit-learn (which the reference
names beside GPy at "Kernel Selection/README.md":9) in tests/test_oracle_gp.py, plus
finite-difference checks of every gradient.

Call sites restated (all in /root/reference/core_navigation/script/gp_slip_node.py):
  :19-30  data prep, 90 % train split                 -> split_train()
  :31     kernel = RBF(1) * Brownian(1)               -> KernelExpr("rbf*brownian")
  :35     GPy.models.GPRegression(x, y, kernel)       -> inference()   (ExactGaussianInference)
  :36     m.optimize()                                -> optimize()    (paramz L-BFGS-B on softplus)
  :45     X_ = arange(X.min(), X.max()+600, 1)        -> prediction_grid()
  :47-50  m.predict([[x]]) per point                  -> predict() / predict_pointwise()
  :59-61  mean[n:], 2*sqrt(var[n:])                   -> gp_slip_callback()

GPy semantics restated (upstream GPy 1.9.x):
  * kern/src/stationary.py  Stationary._unscaled_dist: r^2 by the expanded form
    -2 x x' + (x^2 + x'^2), diagonal forced to 0 when X2 is None, clipped at 0, then sqrt, / lengthscale.
  * kern/src/rbf.py, stationary.py (Matern32, Matern52, RatQuad), standard_periodic.py (StdPeriodic),
    brownian.py, linear.py, static.py (Bias, White), add.py, prod.py: K, Kdiag, update_gradients_full.
  * inference/latent_function_inference/exact_gaussian_inference.py: Ky = K + (sigma_n^2 + 1e-8) I,
    pdinv (dpotrf / dpotri), dpotrs, LML = 0.5(-N log 2pi - logdet - y'alpha), dL_dK = 0.5(alpha alpha' - Ky^-1),
    dL_dthetaL = trace(dL_dK).
  * inference/latent_function_inference/posterior.py PosteriorExact._raw_predict: mu = Kx' alpha,
    tmp = dtrtrs(L, Kx), var = Kxx - sum(tmp^2), clipped at 1e-15; likelihoods/gaussian.py
    predictive_values adds sigma_n^2.
  * util/linalg.py jitchol: on dpotrf failure retry with jitter mean(diag)*1e-6, x10 per try, 5 tries.
  * paramz optimization "lbfgsb" -> scipy.optimize.fmin_l_bfgs_b(maxfun=1000) on Logexp-transformed
    positives: theta = log(1 + exp(x)), gradient factor 1 - exp(-theta).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import lapack

LOG_2_PI = math.log(2.0 * math.pi)
JITTER = 1e-8           # exact_gaussian_inference.py: diag.add(Ky, variance + 1e-8)
VAR_FLOOR = 1e-15       # posterior.py: np.clip(var, 1e-15, np.inf)

# ----------------------------------------------------------------------------------------------
# kernel expressions
# ----------------------------------------------------------------------------------------------
# leaf name -> (opcode, number of hyper-parameters, GPy parameter order)
LEAVES = {
    "rbf": (1, 2, ("variance", "lengthscale")),
    "mat32": (2, 2, ("variance", "lengthscale")),
    "mat52": (3, 2, ("variance", "lengthscale")),
    "ratquad": (4, 3, ("variance", "lengthscale", "power")),
    "stdperiodic": (5, 3, ("variance", "period", "lengthscale")),
    "brownian": (6, 1, ("variance",)),
    "linear": (7, 1, ("variance",)),
    "bias": (8, 1, ("variance",)),
    "white": (9, 1, ("variance",)),
}
ALIASES = {"se": "rbf", "matern32": "mat32", "matern52": "mat52", "rq": "ratquad",
           "periodic": "stdperiodic", "per": "stdperiodic", "const": "bias"}
OP_ADD, OP_MUL = 16, 17
OPNAME = {v[0]: k for k, v in LEAVES.items()}


@dataclass
class KernelExpr:
    """A composite covariance as a postfix program over LEAVES with + and * (GPy Add / Prod).

    `KernelExpr("rbf*brownian")` is the deployed kernel (gp_slip_node.py:31); "rbf+stdperiodic" is the
    SE+periodic candidate of "Kernel Selection/docs/allkernels.jpg".  theta is the concatenation of the
    leaf hyper-parameters in program (left-to-right) order; the Gaussian noise variance is kept separate.
    """
    text: str

    def __post_init__(self):
        self.program: List[int] = _to_postfix(self.text)
        self.n_params = sum(LEAVES[OPNAME[op]][1] for op in self.program if op < OP_ADD)

    def param_names(self) -> List[str]:
        out = []
        k = 0
        for op in self.program:
            if op < OP_ADD:
                nm = OPNAME[op]
                out += [f"{nm}{k}.{p}" for p in LEAVES[nm][2]]
                k += 1
        return out


def _to_postfix(text: str) -> List[int]:
    toks: List[str] = []
    cur = ""
    for ch in text.replace(" ", "").lower():
        if ch in "+*()":
            if cur:
                toks.append(cur)
                cur = ""
            toks.append(ch)
        else:
            cur += ch
    if cur:
        toks.append(cur)
    out: List[int] = []
    stack: List[str] = []
    prec = {"+": 1, "*": 2}
    for t in toks:
        if t in prec:
            while stack and stack[-1] in prec and prec[stack[-1]] >= prec[t]:
                out.append(OP_ADD if stack.pop() == "+" else OP_MUL)
            stack.append(t)
        elif t == "(":
            stack.append(t)
        elif t == ")":
            while stack[-1] != "(":
                out.append(OP_ADD if stack.pop() == "+" else OP_MUL)
            stack.pop()
        else:
            t = ALIASES.get(t, t)
            if t not in LEAVES:
                raise ValueError(f"unknown kernel family {t!r}")
            out.append(LEAVES[t][0])
    while stack:
        out.append(OP_ADD if stack.pop() == "+" else OP_MUL)
    return out


# ----------------------------------------------------------------------------------------------
# leaf covariances (1-D input), GPy forms
# ----------------------------------------------------------------------------------------------
def _r2_expanded(x: np.ndarray, x2: Optional[np.ndarray]) -> np.ndarray:
    """GPy Stationary._unscaled_dist squared, before the lengthscale: expanded form, NOT (x-x')^2."""
    if x2 is None:
        xsq = np.square(x)
        r2 = -2.0 * np.multiply.outer(x, x) + (xsq[:, None] + xsq[None, :])
        np.fill_diagonal(r2, 0.0)
        return np.clip(r2, 0.0, np.inf)
    x1sq = np.square(x)
    x2sq = np.square(x2)
    r2 = -2.0 * np.multiply.outer(x, x2) + (x1sq[:, None] + x2sq[None, :])
    return np.clip(r2, 0.0, np.inf)


def _leaf_K(op: int, th: Sequence[float], x: np.ndarray, x2: Optional[np.ndarray]):
    """Return (K, [dK/dtheta_j ...]) for one leaf.  x2 None means the symmetric K(X, X)."""
    nm = OPNAME[op]
    xx2 = x if x2 is None else x2
    if nm in ("rbf", "mat32", "mat52", "ratquad"):
        var, ls = th[0], th[1]
        r = np.sqrt(_r2_expanded(x, x2)) / ls
        if nm == "rbf":
            kr = np.exp(-0.5 * r ** 2)
            dk_dr = -r * kr
        elif nm == "mat32":
            s3 = np.sqrt(3.0)
            e = np.exp(-s3 * r)
            kr = (1.0 + s3 * r) * e
            dk_dr = -3.0 * r * e
        elif nm == "mat52":
            s5 = np.sqrt(5.0)
            e = np.exp(-s5 * r)
            kr = (1.0 + s5 * r + 5.0 / 3.0 * r ** 2) * e
            dk_dr = (10.0 / 3.0 * r - 5.0 * r - 5.0 * s5 / 3.0 * r ** 2) * e
        else:
            a = th[2]
            l1p = np.log1p(np.square(r) / 2.0)
            kr = np.exp(-a * l1p)
            dk_dr = -a * r * np.exp(-(a + 1.0) * l1p)
        K = var * kr
        grads = [kr, -(var * dk_dr) * r / ls]
        if nm == "ratquad":
            grads.append(-K * l1p)
        return K, grads
    if nm == "stdperiodic":
        var, per, ls = th
        base = np.pi * (x[:, None] - xx2[None, :]) / per
        sb = np.sin(base)
        ed = np.exp(-0.5 * np.square(sb / ls))
        K = var * ed
        dwl = var * (1.0 / ls ** 2) * sb * np.cos(base) * (base / per)
        dl = var * np.square(sb) / ls ** 3
        return K, [ed, dwl * ed, dl * ed]
    if nm == "brownian":
        var = th[0]
        a, b = x[:, None], xx2[None, :]
        base = np.where(np.sign(a) == np.sign(b), np.fmin(np.abs(a), np.abs(b)), 0.0)
        return var * base, [base]
    if nm == "linear":
        base = np.multiply.outer(x, xx2)
        return th[0] * base, [base]
    if nm == "bias":
        base = np.ones((x.size, xx2.size))
        return th[0] * base, [base]
    if nm == "white":
        base = np.eye(x.size) if x2 is None else np.zeros((x.size, xx2.size))
        return th[0] * base, [base]
    raise AssertionError(nm)


def _leaf_Kdiag(op: int, th: Sequence[float], x: np.ndarray) -> np.ndarray:
    nm = OPNAME[op]
    if nm == "brownian":
        return th[0] * np.abs(x)
    if nm == "linear":
        return th[0] * np.square(x)
    return th[0] * np.ones_like(x)


def kernel_K(expr: KernelExpr, theta: Sequence[float], x, x2=None, want_grads=False):
    """K(X, X2) of the composite; with want_grads also the list of dK/dtheta_j matrices (Add / Prod rules)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    x2 = None if x2 is None else np.asarray(x2, dtype=np.float64).ravel()
    stack = []
    p = 0
    for op in expr.program:
        if op < OP_ADD:
            n = LEAVES[OPNAME[op]][1]
            K, g = _leaf_K(op, theta[p:p + n], x, x2)
            stack.append((K, {p + j: g[j] for j in range(n)}))
            p += n
        else:
            Kb, gb = stack.pop()
            Ka, ga = stack.pop()
            if op == OP_ADD:
                K = Ka + Kb
                g = {**ga, **gb}
            else:
                K = Ka * Kb
                g = {**{k: v * Kb for k, v in ga.items()}, **{k: v * Ka for k, v in gb.items()}}
            stack.append((K, g))
    K, g = stack.pop()
    if want_grads:
        return K, [g[j] for j in range(expr.n_params)]
    return K


def kernel_Kdiag(expr: KernelExpr, theta: Sequence[float], x) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64).ravel()
    stack = []
    p = 0
    for op in expr.program:
        if op < OP_ADD:
            n = LEAVES[OPNAME[op]][1]
            stack.append(_leaf_Kdiag(op, theta[p:p + n], x))
            p += n
        else:
            b = stack.pop()
            a = stack.pop()
            stack.append(a + b if op == OP_ADD else a * b)
    return stack.pop()


# ----------------------------------------------------------------------------------------------
# exact inference, prediction
# ----------------------------------------------------------------------------------------------
class NotPositiveDefinite(Exception):
    pass


def jitchol(A: np.ndarray, maxtries: int = 5) -> Tuple[np.ndarray, float]:
    """GPy util/linalg.py jitchol: dpotrf, then a x10 jitter ladder from mean(diag)*1e-6.  Returns (L, jitter used)."""
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return np.tril(L), 0.0
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise NotPositiveDefinite("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        L, info = lapack.dpotrf(A + np.eye(A.shape[0]) * jitter, lower=1)
        if info == 0:
            return np.tril(L), jitter
        jitter *= 10.0
    raise NotPositiveDefinite("not positive definite, even with jitter.")


@dataclass
class Inference:
    L: np.ndarray
    alpha: np.ndarray
    lml: float
    logdet: float
    Kinv: Optional[np.ndarray]
    grad: Optional[np.ndarray]      # d LML / d [theta..., noise]
    jitter: float


def inference(expr: KernelExpr, theta, noise: float, x, y, want_grad: bool = False) -> Inference:
    """ExactGaussianInference.inference at fixed hyper-parameters (gp_slip_node.py:35)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    n = x.size
    if want_grad:
        K, dKs = kernel_K(expr, theta, x, None, want_grads=True)
    else:
        K, dKs = kernel_K(expr, theta, x, None), None
    Ky = K.copy()
    Ky[np.diag_indices(n)] += noise + JITTER
    L, jit = jitchol(Ky)
    alpha, info = lapack.dpotrs(L, y.reshape(-1, 1), lower=1)
    assert info == 0
    alpha = alpha.ravel()
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    lml = 0.5 * (-n * LOG_2_PI - logdet - float(y @ alpha))
    Kinv = grad = None
    if want_grad:
        Ki, info = lapack.dpotri(L, lower=1)
        assert info == 0
        Kinv = np.tril(Ki) + np.tril(Ki, -1).T
        dL_dK = 0.5 * (np.multiply.outer(alpha, alpha) - Kinv)
        grad = np.array([np.sum(dL_dK * dK) for dK in dKs] + [np.trace(dL_dK)])
    return Inference(L, alpha, lml, logdet, Kinv, grad, jit)


def predict(expr: KernelExpr, theta, noise: float, x, y, xnew, inf: Optional[Inference] = None,
            include_likelihood: bool = True):
    """m.predict(Xnew) with all M points at once (vectorised dtrtrs).  Returns (mean[M], var[M])."""
    x = np.asarray(x, dtype=np.float64).ravel()
    xnew = np.asarray(xnew, dtype=np.float64).ravel()
    if inf is None:
        inf = inference(expr, theta, noise, x, y)
    Kx = kernel_K(expr, theta, x, xnew)
    mu = Kx.T @ inf.alpha
    Kxx = kernel_Kdiag(expr, theta, xnew)
    tmp, info = lapack.dtrtrs(inf.L, Kx, lower=1)
    assert info == 0
    var = np.clip(Kxx - np.square(tmp).sum(0), VAR_FLOOR, np.inf)
    if include_likelihood:
        var = var + noise
    return mu, var


def predict_pointwise(expr: KernelExpr, theta, noise: float, x, y, xnew, inf: Optional[Inference] = None):
    """The reference's own loop shape: one m.predict([[x]]) per test point (gp_slip_node.py:47-50)."""
    xnew = np.asarray(xnew, dtype=np.float64).ravel()
    if inf is None:
        inf = inference(expr, theta, noise, x, y)
    mu = np.empty(xnew.size)
    var = np.empty(xnew.size)
    for k in range(xnew.size):
        m, v = predict(expr, theta, noise, x, y, xnew[k:k + 1], inf)
        mu[k], var[k] = m[0], v[0]
    return mu, var


# ----------------------------------------------------------------------------------------------
# hyper-parameter fit (gp_slip_node.py:36) - paramz lbfgsb on Logexp (softplus) transformed positives
# ----------------------------------------------------------------------------------------------
_LIM = 36.0  # paramz transformations.py _lim_val


def softplus(z):
    z = np.asarray(z, dtype=np.float64)
    return np.where(z > _LIM, z, np.log1p(np.exp(np.clip(z, -np.inf, _LIM))))


def softplus_inv(t):
    t = np.asarray(t, dtype=np.float64)
    return np.where(t > _LIM, t, np.log(np.expm1(t)))


def softplus_gradfactor(t):
    t = np.asarray(t, dtype=np.float64)
    return np.where(t > _LIM, 1.0, -np.expm1(-t))


def optimize(expr: KernelExpr, x, y, theta0=None, noise0: float = 1.0, max_iters: int = 1000):
    """m.optimize(): minimise -LML over softplus-transformed [theta, noise] with scipy L-BFGS-B.

    Returns (theta, noise, lml, n_evals).  Optimised hyper-parameters are NOT expected to be 1e-9-reproducible
    across implementations; parity of the fit is "reaches the same LML".
    """
    from scipy.optimize import fmin_l_bfgs_b
    p0 = np.ones(expr.n_params + 1) if theta0 is None else np.append(np.asarray(theta0, float), noise0)
    z0 = softplus_inv(p0)
    nev = [0]

    def f_fp(z):
        p = softplus(z)
        nev[0] += 1
        try:
            inf = inference(expr, p[:-1], p[-1], x, y, want_grad=True)
        except NotPositiveDefinite:
            return 1e300, np.zeros_like(z)
        return -inf.lml, -inf.grad * softplus_gradfactor(p)

    z, fval, _ = fmin_l_bfgs_b(f_fp, z0, maxfun=max_iters, maxiter=max_iters)
    p = softplus(z)
    return p[:-1], float(p[-1]), -float(fval), nev[0]


# ----------------------------------------------------------------------------------------------
# the node callback (gp_slip_node.py:16-63)
# ----------------------------------------------------------------------------------------------
TRAIN_FRACTION = 0.9   # gp_slip_node.py:27
HORIZON = 600          # gp_slip_node.py:45
GRID_STEP = 1.0        # gp_slip_node.py:45


def split_train(time_array, slip_array):
    """gp_slip_node.py:19-30: first int(0.9 n) samples train the GP."""
    X = np.asarray(time_array, dtype=np.float64)
    Y = np.asarray(slip_array, dtype=np.float64)
    ntr = int(TRAIN_FRACTION * len(X))
    return X[:ntr], Y[:ntr]


def prediction_grid(time_array):
    """gp_slip_node.py:45: np.arange(X.min(), X.max() + 600, 1)."""
    X = np.asarray(time_array, dtype=np.float64)
    return np.arange(X.min(), X.max() + HORIZON, GRID_STEP)


def gp_slip_callback(time_array, slip_array, expr: KernelExpr = None, theta=None, noise=None):
    """Restatement of gp_slip_node.callback.  With theta/noise None the hypers are fitted (m.optimize()).

    Returns (mean, sigma) exactly as published in core_nav/GP_Output: mean = means[len(X):],
    sigma = 2*sqrt(variances[len(X):]) (gp_slip_node.py:59-61) - sigma is TWO standard deviations and
    includes the noise variance.
    """
    expr = expr or KernelExpr("rbf*brownian")
    xtr, ytr = split_train(time_array, slip_array)
    if theta is None:
        theta, noise, _, _ = optimize(expr, xtr, ytr)
    grid = prediction_grid(time_array)
    mu, var = predict(expr, theta, noise, xtr, ytr, grid)
    n = len(time_array)
    return mu[n:], 2.0 * np.sqrt(var[n:])
