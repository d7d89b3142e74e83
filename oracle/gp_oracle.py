"""Oracle for one GP region (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates `fit_gp_spp` (/root/reference/gapro/gaussian_process_utils.py:382-445)
and `GPClassificationModel` (:11-25).  All arithmetic of that function lives in
gpytorch (absent here; restated at 1.8.1-style constants):

  * whitened `VariationalStrategy` with `CholeskyVariationalDistribution(M)`,
    inducing points initialised at the M training rows and trainable;
  * `ConstantMean`, `ScaleKernel(RBFKernel)` with softplus-constrained raw
    parameters initialised at 0;
  * `BernoulliLikelihood` (probit), 20-point Gauss-Hermite expected log-prob,
    labels used as -1/+1 directly;
  * `VariationalELBO(num_data=M)`, `torch.optim.Adam(lr=0.1)`, 50 steps;
  * K_ZZ jitter 1e-4 (`variational_cholesky_jitter`, float), K_XX jitter 1e-4,
    `min_variance` 1e-6, `mean_init_std` 1e-3.

Two implementations:

  `fit_region_autograd`   torch autograd through the gpytorch-shaped forward
                          (quadratic-expansion distances, Cholesky +
                          triangular solve).  Precision policy "fp64" (all
                          float64 — the parity target of the CUDA path) or
                          "gpytorch" (float32 everywhere except the float64
                          Cholesky/solve — what the reference executes).
  `fit_region_manual`     numpy float64, hand-derived gradient (SURVEY.md
                          §8a-C), explicit inverse of the Cholesky factor and
                          direct-difference distances: the algorithm the CUDA
                          kernels implement, step for step.

The random initial variational mean (gpytorch draws it from the unseeded global
RNG) is an explicit input `init_noise` (standard-normal draws, one per training
row); the mean is initialised to 1e-3 * init_noise.
"""
from __future__ import annotations

import math

import numpy as np
import torch

N_GH = 20
_GH_T, _GH_W = np.polynomial.hermite.hermgauss(N_GH)
MIN_VARIANCE = 1e-6
MEAN_INIT_STD = 1e-3
ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-8


# --------------------------------------------------------------------------- #
# gpytorch-shaped forward (torch, autograd)
# --------------------------------------------------------------------------- #
def _sq_dist(x1, x2):
    """gpytorch `Distance._sq_dist` (kernels/kernel.py): centre on mean(x1),
    quadratic expansion with inner dimension D+2, clamp at 0."""
    adj = x1.mean(-2, keepdim=True)
    x1 = x1 - adj
    x2 = x2 - adj
    x1_norm = x1.pow(2).sum(-1, keepdim=True)
    x2_norm = x2.pow(2).sum(-1, keepdim=True)
    x1_ = torch.cat([-2.0 * x1, x1_norm, torch.ones_like(x1_norm)], -1)
    x2_ = torch.cat([x2, torch.ones_like(x2_norm), x2_norm], -1)
    return x1_.matmul(x2_.transpose(-2, -1)).clamp_min(0)


def _psd_safe_cholesky(K):
    """gpytorch `psd_safe_cholesky`: retry with +1e-8*10^k on the diagonal (float64)."""
    L, info = torch.linalg.cholesky_ex(K)
    if not bool(info.any()):
        return L
    jitter_prev = 0.0
    Kp = K.clone()
    for i in range(3):
        jitter_new = 1e-8 * (10 ** i)
        Kp.diagonal().add_(jitter_new - jitter_prev)
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Kp)
        if not bool(info.any()):
            return L
    raise RuntimeError("NotPSDError: matrix not positive definite after jitter retries")


def _forward(params, X, jitter_zz, jitter_xx, policy):
    """q(f) at rows X: returns (mu, var) with var already clamped at MIN_VARIANCE."""
    Z, m, Lq, c, rho_s, rho_l = params
    ell = torch.nn.functional.softplus(rho_l)
    s = torch.nn.functional.softplus(rho_s)
    Zs, Xs = Z / ell, X / ell
    Kzz = s * torch.exp(-0.5 * _sq_dist(Zs, Zs))
    Kzz = Kzz + jitter_zz * torch.eye(Z.shape[0], dtype=Z.dtype)
    Kzx = s * torch.exp(-0.5 * _sq_dist(Zs, Xs))
    L = _psd_safe_cholesky(Kzz.double())
    A = torch.linalg.solve_triangular(L, Kzx.double(), upper=False)
    if policy == "gpytorch":
        A = A.float()
    mu = A.transpose(0, 1) @ m + c
    T = Lq * torch.ones_like(Lq).tril(0)
    # K_XX diag + jitter + diag(A^T (T T^T - I) A)
    TB = T @ (T.transpose(0, 1) @ A) - A
    v = s + jitter_xx + (A * TB).sum(0)
    var = v.clamp_min(MIN_VARIANCE)
    return mu, var, T


def _neg_elbo(params, X, y, jitter_zz, jitter_xx, policy):
    mu, var, T = _forward(params, X, jitter_zz, jitter_xx, policy)
    dt = mu.dtype
    t = torch.as_tensor(_GH_T, dtype=dt)
    w = torch.as_tensor(_GH_W, dtype=dt)
    locs = torch.sqrt(2.0 * var)[None, :] * t[:, None] + mu[None, :]
    logp = torch.special.log_ndtr(locs * y[None, :])
    ell_i = (1.0 / math.sqrt(math.pi)) * (logp * w[:, None]).sum(0)
    n = X.shape[0]
    m = params[1]
    kl = 0.5 * ((m * m).sum() + (T * T).sum() - T.diagonal().pow(2).log().sum() - m.shape[0])
    return -(ell_i.sum() / n) + kl / n


def fit_region_autograd(train_x, n_b1, test_x, init_noise, iters=50, lr=0.1,
                        jitter_zz=1e-4, jitter_xx=1e-4, policy="fp64", return_params=False):
    """train_x (M,D): rows of box b1 first (label -1), then box b2 (label +1).
    Returns dict(prob, conf, label, mu, var) as float32/bool numpy arrays of
    length len(test_x) (cast exactly as the product casts them)."""
    dt = torch.float64 if policy == "fp64" else torch.float32
    X = torch.as_tensor(np.asarray(train_x), dtype=dt)
    Xt = torch.as_tensor(np.asarray(test_x), dtype=dt)
    M = X.shape[0]
    y = torch.cat([-torch.ones(n_b1, dtype=dt), torch.ones(M - n_b1, dtype=dt)])
    noise = torch.as_tensor(np.asarray(init_noise), dtype=torch.float64)
    Z = X.clone().requires_grad_(True)
    m = (MEAN_INIT_STD * noise).to(dt).requires_grad_(True)
    Lq = torch.eye(M, dtype=dt).requires_grad_(True)
    c = torch.zeros((), dtype=dt, requires_grad=True)
    rho_s = torch.zeros((), dtype=dt, requires_grad=True)
    rho_l = torch.zeros((), dtype=dt, requires_grad=True)
    params = [Z, m, Lq, c, rho_s, rho_l]
    opt = torch.optim.Adam(params, lr=lr)
    for _ in range(iters):
        loss = _neg_elbo(params, X, y, jitter_zz, jitter_xx, policy)
        opt.zero_grad()
        loss.backward()
        opt.step()
    with torch.no_grad():
        mu, var, _ = _forward(params, Xt, jitter_zz, jitter_xx, policy)
    out = finish_prediction(mu.double().numpy(), var.double().numpy())
    if return_params:
        out["params"] = [p.detach().double().numpy() for p in params]
    return out


def finish_prediction(mu64, var64):
    """`likelihood(f_pred).mean` = Phi(mu/sqrt(1+var)), then the float32 casts and
    the label/confidence rule of gaussian_process_utils.py:432-438."""
    link = mu64 / np.sqrt(1.0 + var64)
    from scipy.special import erfc
    p64 = 0.5 * erfc(-link / math.sqrt(2.0))
    p32 = p64.astype(np.float32)
    label = p32 >= np.float32(0.5)
    conf = np.where(label, p32, np.float32(1.0) - p32).astype(np.float32)
    return dict(prob=p32, conf=conf, label=label, mu=mu64.astype(np.float32),
                var=var64.astype(np.float32), mu64=mu64, var64=var64, prob64=p64)


# --------------------------------------------------------------------------- #
# hand-derived gradient, numpy float64 — the algorithm of the CUDA path
# --------------------------------------------------------------------------- #
def _softplus(x):
    return math.log1p(math.exp(-abs(x))) + max(x, 0.0)


def _sigmoid(x):
    return 1.0 / (1.0 + math.exp(-x))


def _rbf(Za, Xb, ell, s):
    d = Za[:, None, :] - Xb[None, :, :]
    r2 = (d * d).sum(-1) / (ell * ell)
    return s * np.exp(-0.5 * r2), r2


def hazard(z):
    """phi(z)/Phi(z) = sqrt(2/pi) / erfcx(-z/sqrt 2): derivative of log Phi."""
    from scipy.special import erfcx
    with np.errstate(over="ignore"):
        return math.sqrt(2.0 / math.pi) / erfcx(-z / math.sqrt(2.0))


def manual_grads(Z, m, T, c, rho_s, rho_l, X, y, jitter_zz, jitter_xx, return_internals=False):
    """One forward+backward of -ELBO with the hand-derived gradient.
    T is the already-masked lower-triangular factor.  Returns gradients
    (gZ, gm, gT, gc, grho_s, grho_l)."""
    from scipy.linalg import solve_triangular
    M, N = Z.shape[0], X.shape[0]
    ell, s = _softplus(rho_l), _softplus(rho_s)
    Kzz0, r2zz = _rbf(Z, Z, ell, s)
    Kzx, r2zx = _rbf(Z, X, ell, s)
    Kzz = Kzz0 + jitter_zz * np.eye(M)
    L = np.linalg.cholesky(Kzz)
    Linv = solve_triangular(L, np.eye(M), lower=True)
    A = Linv @ Kzx
    mu = A.T @ m + c
    B = T.T @ A
    v = s + jitter_xx + (B * B).sum(0) - (A * A).sum(0)
    clamped = v < MIN_VARIANCE
    var = np.where(clamped, MIN_VARIANCE, v)
    # Gauss-Hermite expected log-prob gradients
    sd = np.sqrt(2.0 * var)
    zk = y[None, :] * (sd[None, :] * _GH_T[:, None] + mu[None, :])
    h = hazard(zk)
    pref = -(1.0 / N) / math.sqrt(math.pi)
    g_mu = pref * (y[None, :] * h * _GH_W[:, None]).sum(0)
    g_v = pref * (y[None, :] * h * (_GH_W * _GH_T)[:, None]).sum(0) / sd
    g_v = np.where(clamped, 0.0, g_v)
    gc = g_mu.sum()
    gm = A @ g_mu + m / N
    gT = np.tril(2.0 * (A * g_v[None, :]) @ B.T + (T - np.diag(1.0 / np.diag(T))) / N)
    G_A = np.outer(m, g_mu) + 2.0 * (T @ B - A) * g_v[None, :]
    g_s = g_v.sum()
    G_C = Linv.T @ G_A                       # dLoss/dK_zx
    G_L = -np.tril(G_C @ A.T)
    P = np.tril(L.T @ G_L)
    P[np.diag_indices(M)] *= 0.5
    symP = 0.5 * (P + P.T)
    G_K = Linv.T @ symP @ Linv               # dLoss/dK_zz (symmetric)
    Ezz, Ezx = Kzz0 / s, Kzx / s
    g_s += (G_K * Ezz).sum() + (G_C * Ezx).sum()
    Gr_zz = -0.5 * G_K * Kzz0
    Gr_zx = -0.5 * G_C * Kzx
    g_l = (Gr_zz * (-2.0 * r2zz / ell)).sum() + (Gr_zx * (-2.0 * r2zx / ell)).sum()
    W = Gr_zz + Gr_zz.T
    gZ = (2.0 / (ell * ell)) * (
        W.sum(1)[:, None] * Z - W @ Z + Gr_zx.sum(1)[:, None] * Z - Gr_zx @ X
    )
    grads = (gZ, gm, gT, gc, g_s * _sigmoid(rho_s), g_l * _sigmoid(rho_l))
    if return_internals:
        return grads, dict(L=L, Linv=Linv, A=A, B=B, mu=mu, var=var, g_mu=g_mu, g_v=g_v,
                           G_A=G_A, G_C=G_C, G_L=G_L, symP=symP, G_K=G_K, Kzx=Kzx)
    return grads


class _Adam:
    """torch.optim.Adam (no amsgrad / weight decay) for a list of numpy arrays."""

    def __init__(self, params, lr):
        self.params, self.lr, self.t = params, lr, 0
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]

    def step(self, grads):
        self.t += 1
        bc1 = 1.0 - ADAM_BETA1 ** self.t
        bc2s = math.sqrt(1.0 - ADAM_BETA2 ** self.t)
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m *= ADAM_BETA1
            m += (1.0 - ADAM_BETA1) * g
            v *= ADAM_BETA2
            v += (1.0 - ADAM_BETA2) * g * g
            p -= (self.lr / bc1) * m / (np.sqrt(v) / bc2s + ADAM_EPS)


def predict_manual(Z, m, T, c, rho_s, rho_l, Xt, jitter_zz, jitter_xx):
    from scipy.linalg import solve_triangular
    M = Z.shape[0]
    ell, s = _softplus(rho_l), _softplus(rho_s)
    Kzz, _ = _rbf(Z, Z, ell, s)
    Kzx, _ = _rbf(Z, Xt, ell, s)
    L = np.linalg.cholesky(Kzz + jitter_zz * np.eye(M))
    A = solve_triangular(L, np.eye(M), lower=True) @ Kzx
    B = T.T @ A
    mu = A.T @ m + c
    v = s + jitter_xx + (B * B).sum(0) - (A * A).sum(0)
    return mu, np.maximum(v, MIN_VARIANCE)


def fit_region_manual(train_x, n_b1, test_x, init_noise, iters=50, lr=0.1,
                      jitter_zz=1e-4, jitter_xx=1e-4, return_params=False):
    X = np.asarray(train_x, dtype=np.float64)
    Xt = np.asarray(test_x, dtype=np.float64)
    M = X.shape[0]
    y = np.concatenate([-np.ones(n_b1), np.ones(M - n_b1)])
    Z = X.copy()
    m = MEAN_INIT_STD * np.asarray(init_noise, dtype=np.float64)
    T = np.eye(M)
    sc = [np.zeros(()), np.zeros(()), np.zeros(())]   # c, rho_s, rho_l
    params = [Z, m, T] + sc
    opt = _Adam(params, lr)
    for _ in range(iters):
        g = manual_grads(Z, m, T, float(sc[0]), float(sc[1]), float(sc[2]), X, y, jitter_zz, jitter_xx)
        opt.step([np.asarray(x) for x in g])
    mu, var = predict_manual(Z, m, T, float(sc[0]), float(sc[1]), float(sc[2]), Xt, jitter_zz, jitter_xx)
    out = finish_prediction(mu, var)
    if return_params:
        out["params"] = [np.asarray(p, dtype=np.float64) for p in params]
    return out


def fit_gp_points_oracle(coords_float, feats, spp, b1_inds, b2_inds, intersect_inds, init_noise, training_iter=50,
                         npoint_nearest=800, spp_pool=True):
    """Restates the point-level variant fit_gp (/root/reference/gapro/gaussian_process_utils.py:28-116)
    on top of fit_region_autograd (fp64 policy).  Returns (probs, conf, labels, bernoulli_variance)."""
    from .gen_ps_oracle import scatter_sum_index_order
    coords = np.asarray(coords_float, dtype=np.float64)
    feats = np.asarray(feats, dtype=np.float32)
    spp = np.asarray(spp)

    def pooled(idx):
        _, dense = np.unique(spp[idx], return_inverse=True)
        dense = dense.reshape(-1)
        s, cnt = scatter_sum_index_order(feats[idx], dense, int(dense.max()) + 1)
        return s / np.maximum(cnt, 1).astype(np.float32)[:, None]

    def nearest(idx):
        if len(idx) <= npoint_nearest:
            return feats[idx]
        c = coords[intersect_inds].mean(0)
        d = ((coords[idx] - c[None, :]) ** 2).sum(1)
        return feats[idx][np.argsort(d, kind="stable")[:npoint_nearest]]

    f1, f2 = (pooled(b1_inds), pooled(b2_inds)) if spp_pool else (nearest(b1_inds), nearest(b2_inds))
    r = fit_region_autograd(np.concatenate([f1, f2]), len(f1), feats[intersect_inds], init_noise, iters=training_iter)
    p = r["prob"]
    return p, r["conf"], r["label"], (p * (np.float32(1.0) - p)).astype(np.float32)


def fit_gp_ensemble_oracle(coords_float, feats, spp, b1_inds, b2_inds, intersect_inds, channel_dims, init_noise,
                           training_iter=50, npoint_nearest=800, spp_pool=True):
    """Restates fit_gp_ensemble (/root/reference/gapro/gaussian_process_utils.py:119-251), fp64 policy per member.
    Returns (probs, labels, variance) per intersection point."""
    from .gen_ps_oracle import scatter_sum_index_order
    coords = np.asarray(coords_float, dtype=np.float64)
    feats = np.asarray(feats, dtype=np.float32)
    spp = np.asarray(spp)
    c = coords[intersect_inds].mean(0)

    def nearest(idx):
        if len(idx) <= npoint_nearest:
            return idx
        d = ((coords[idx] - c[None, :]) ** 2).sum(1)
        return idx[np.argsort(d, kind="stable")[:npoint_nearest]]

    def pooled(idx):
        _, dense = np.unique(spp[idx], return_inverse=True)
        dense = dense.reshape(-1)
        s, cnt = scatter_sum_index_order(feats[idx], dense, int(dense.max()) + 1)
        return s / np.maximum(cnt, 1).astype(np.float32)[:, None], dense

    i1, i2 = nearest(np.asarray(b1_inds)), nearest(np.asarray(b2_inds))
    if spp_pool:
        f1, f2 = pooled(i1)[0], pooled(i2)[0]
        ft, inverse = pooled(np.asarray(intersect_inds))
    else:
        f1, f2, ft, inverse = feats[i1], feats[i2], feats[intersect_inds], None
    score = np.zeros((len(ft), 2), dtype=np.float32)
    variance = np.zeros(len(ft), dtype=np.float32)
    for g in range(len(channel_dims) - 1):
        a, b = channel_dims[g], channel_dims[g + 1]
        r = fit_region_autograd(np.concatenate([f1[:, a:b], f2[:, a:b]]), len(f1), ft[:, a:b], init_noise[g],
                                iters=training_iter)
        p, lab = r["prob"], r["label"]
        score[:, 1] += np.where(lab, p, np.float32(1) - p)
        score[:, 0] += np.where(lab, np.float32(1) - p, p)
        variance += p * (np.float32(1) - p)
    labels = score.argmax(1)
    probs = score.max(1)
    if inverse is not None:
        probs, labels, variance = probs[inverse], labels[inverse], variance[inverse]
    return probs, labels, variance
