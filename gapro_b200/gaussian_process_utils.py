"""Host-side mirror of /root/reference/gapro/gaussian_process_utils.py for the GP region fit.

`fit_gp_spp` keeps the reference signature and return tuple; `fit_gp_regions` is the batched
form the pipeline uses.  The model the reference builds with gpytorch
(GPClassificationModel, gaussian_process_utils.py:11-25: whitened VariationalStrategy with a
CholeskyVariationalDistribution over M inducing points initialised at the training rows,
ConstantMean, ScaleKernel(RBFKernel), BernoulliLikelihood, VariationalELBO, Adam lr 0.1) is
implemented in libgapro_b200.so (csrc/gp_fit.cu); nothing here does arithmetic.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["fit_gp_spp", "fit_gp", "fit_gp_ensemble", "fit_gp_regions"]


def fit_gp_regions(feats_spp, train_lists, n_b1, test_lists, init_noise=None, training_iter=50, lr=0.1,
                   jitter_zz=1e-4, jitter_xx=1e-4, workspace_bytes=None, return_float64=False):
    """Fit R independent regions in one batched call.

    feats_spp   (S,D) float32 CUDA tensor of pooled superpoint features
    train_lists list of R int tensors/arrays: training rows (box-1 rows first)
    n_b1        list of R ints: how many leading training rows belong to box 1 (label -1)
    test_lists  list of R int tensors/arrays: rows to predict
    init_noise  list of R float arrays of standard-normal draws (None: torch.randn on device)
    Returns a list of R tuples (pred_probs, pred_probs_new, pred_labels, pred_mu, pred_variance)
    (+ (mu64, var64) when return_float64)."""
    if not feats_spp.is_cuda:
        raise _lib.GaproError("fit_gp_regions needs CUDA tensors; gapro_b200 has no CPU fallback")
    lib = _lib.load()
    dev = feats_spp.device
    feats = feats_spp.float().contiguous()
    R = len(train_lists)
    if R == 0:
        return []
    as_np = lambda x: (x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)).astype(np.int32).reshape(-1)
    tr = [as_np(t) for t in train_lists]
    te = [as_np(t) for t in test_lists]
    train_off = np.zeros(R + 1, dtype=np.int32)
    train_off[1:] = np.cumsum([len(t) for t in tr])
    test_off = np.zeros(R + 1, dtype=np.int32)
    test_off[1:] = np.cumsum([len(t) for t in te])
    nb1 = np.asarray(n_b1, dtype=np.int32)
    D = int(feats.shape[1])
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    train_idx = to_dev(np.concatenate(tr))
    test_idx = to_dev(np.concatenate(te))
    if init_noise is None:
        noise = torch.randn(int(train_off[-1]), dtype=torch.float32, device=dev)
    else:
        noise = to_dev(np.concatenate([np.asarray(z, dtype=np.float32).reshape(-1) for z in init_noise]))
        if noise.numel() != int(train_off[-1]):
            raise ValueError("init_noise must hold one draw per training row")
    nt = int(test_off[-1])
    prob = torch.empty(nt, dtype=torch.float32, device=dev)
    conf = torch.empty_like(prob)
    mu = torch.empty_like(prob)
    var = torch.empty_like(prob)
    label = torch.empty(nt, dtype=torch.uint8, device=dev)
    mu64 = torch.empty(nt, dtype=torch.float64, device=dev) if return_float64 else None
    var64 = torch.empty(nt, dtype=torch.float64, device=dev) if return_float64 else None
    status = torch.zeros(R, dtype=torch.int32, device=dev)
    full = lib.gapro_gp_workspace_bytes(R, train_off.ctypes.data, test_off.ctypes.data, D)
    need = lib.gapro_gp_min_workspace_bytes(R, train_off.ctypes.data, test_off.ctypes.data, D)
    nbytes = full if workspace_bytes is None else max(min(full, int(workspace_bytes)), need)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: 0 if t is None else t.data_ptr()
    _lib.check(lib.gapro_gp_fit_batch(feats.data_ptr(), D, R, train_off.ctypes.data, nb1.ctypes.data,
                                      test_off.ctypes.data, train_idx.data_ptr(), test_idx.data_ptr(),
                                      noise.data_ptr(), int(training_iter), float(lr), float(jitter_zz),
                                      float(jitter_xx), prob.data_ptr(), conf.data_ptr(), label.data_ptr(),
                                      mu.data_ptr(), var.data_ptr(), p(mu64), p(var64), status.data_ptr(),
                                      ws.data_ptr(), ws.numel(), stream), "gapro_gp_fit_batch")
    st = status.cpu().numpy()
    fit_gp_regions.last_retries = ((st >> _lib.GP_RETRY_SHIFT) & 0xffff).astype(np.int64)   # psd_safe_cholesky retries
    if np.any(st & _lib.GP_NOT_PSD):
        raise _lib.GaproError("NotPSDError: K_ZZ not positive definite after the jitter retries in region %d"
                              % int(np.flatnonzero(st & 1)[0]))
    if np.any(st & _lib.GP_NAN):
        raise _lib.GaproError("NanError: non-finite GP posterior")
    out = []
    for r in range(R):
        a, b = int(test_off[r]), int(test_off[r + 1])
        item = (prob[a:b], conf[a:b], label[a:b].bool(), mu[a:b], var[a:b])
        if return_float64:
            item = item + (mu64[a:b], var64[a:b])
        out.append(item)
    return out


def fit_gp_spp(coords_float_spp, feats_spp, b1_inds, b2_inds, intersect_inds, training_iter=50, *, init_noise=None):
    """Drop-in for fit_gp_spp (/root/reference/gapro/gaussian_process_utils.py:382-445).

    Trains on feats_spp[b1_inds] (label -1) and feats_spp[b2_inds] (label +1), predicts at
    feats_spp[intersect_inds]; returns (pred_probs, pred_probs_new, pred_labels, pred_mu,
    pred_variance), each of length len(intersect_inds).  `coords_float_spp` is accepted and
    unused, as in the reference (:385-392).  `init_noise` (M standard-normal draws) replaces
    gpytorch's unseeded random initialisation of the variational mean."""
    b1 = b1_inds.reshape(-1)
    b2 = b2_inds.reshape(-1)
    train = torch.cat([b1, b2])
    res = fit_gp_regions(feats_spp, [train], [int(b1.numel())], [intersect_inds],
                         init_noise=None if init_noise is None else [init_noise], training_iter=training_iter)
    return res[0]


def _pool_by_superpoint(feats_rows, spp_rows, return_inverse=False):
    """scatter-mean of `feats_rows` over unique(spp_rows) in sorted-id order, index-ordered float32 sums
    (gapro_densify_spp + gapro_pool_feats): what gaussian_process_utils.py:66-70 does with torch_scatter."""
    lib = _lib.load()
    dev = feats_rows.device
    n, D = int(feats_rows.shape[0]), int(feats_rows.shape[1])
    stream = torch.cuda.current_stream(dev).cuda_stream
    raw = spp_rows.to(dev, torch.int64).contiguous()
    pt_off = np.array([0, n], dtype=np.int64)
    gid = torch.empty(n, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    seg = torch.empty(n + 1, dtype=torch.int32, device=dev)
    spp_off = np.zeros(2, dtype=np.int32)
    ws = torch.empty(lib.gapro_densify_workspace_bytes(n, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.gapro_densify_spp(raw.data_ptr(), pt_off.ctypes.data, 1, gid.data_ptr(), perm.data_ptr(),
                                     seg.data_ptr(), spp_off.ctypes.data, ws.data_ptr(), ws.numel(), stream),
               "gapro_densify_spp")
    S = int(spp_off[1])
    out = torch.empty((S, D), dtype=torch.float32, device=dev)
    src = feats_rows.float().contiguous()
    _lib.check(lib.gapro_pool_feats(src.data_ptr(), perm.data_ptr(), seg.data_ptr(), S, D, out.data_ptr(), stream),
               "gapro_pool_feats")
    return (out, gid.long()) if return_inverse else out


def _nearest_rows(coords_rows, centroid, npoint_nearest):
    """rows kept by the top-k-nearest rule (gaussian_process_utils.py:73-83): all of them up to
    `npoint_nearest`, else the `npoint_nearest` closest to `centroid`, nearest first."""
    n = int(coords_rows.shape[0])
    if n <= npoint_nearest:
        return torch.arange(n, device=coords_rows.device)
    d = ((coords_rows - centroid[None, :]) ** 2).sum(1)
    return torch.topk(d, k=npoint_nearest, largest=False)[1]


def fit_gp(coords_float, feats, spp, b1_inds, b2_inds, intersect_inds, training_iter=50, npoint_nearest=800,
           spp_pool=True, *, init_noise=None):
    """Drop-in for the point-level variant fit_gp (/root/reference/gapro/gaussian_process_utils.py:28-116):
    training rows are the superpoint means of the points of each box (spp_pool=True) or the
    `npoint_nearest` points closest to the centroid of the intersection (spp_pool=False); prediction is
    per intersection POINT.  Returns (pred_probs, pred_probs_new, pred_labels, pred_variance) with
    pred_variance the Bernoulli variance p(1-p), as there (:110)."""
    if not feats.is_cuda:
        raise _lib.GaproError("fit_gp needs CUDA tensors; gapro_b200 has no CPU fallback")
    feats = feats.float()
    b1_feats, b2_feats = feats[b1_inds], feats[b2_inds]
    if spp_pool:
        b1_feats = _pool_by_superpoint(b1_feats, spp[b1_inds])
        b2_feats = _pool_by_superpoint(b2_feats, spp[b2_inds])
    else:
        centroid = coords_float[intersect_inds].mean(0)
        b1_feats = b1_feats[_nearest_rows(coords_float[b1_inds], centroid, npoint_nearest)]
        b2_feats = b2_feats[_nearest_rows(coords_float[b2_inds], centroid, npoint_nearest)]
    n1, n2 = int(b1_feats.shape[0]), int(b2_feats.shape[0])
    test = feats[intersect_inds]
    table = torch.cat([b1_feats, b2_feats, test], 0).contiguous()
    res = fit_gp_regions(table, [np.arange(n1 + n2)], [n1], [np.arange(n1 + n2, n1 + n2 + int(test.shape[0]))],
                         init_noise=None if init_noise is None else [init_noise], training_iter=training_iter)[0]
    probs, conf, labels = res[0], res[1], res[2]
    return probs, conf, labels, probs * (1 - probs)


def fit_gp_ensemble(coords_float, feats, spp, b1_inds, b2_inds, intersect_inds, channel_dims, training_iter=50,
                    npoint_nearest=800, spp_pool=True, *, init_noise=None):
    """Drop-in for fit_gp_ensemble (/root/reference/gapro/gaussian_process_utils.py:119-251): one GP per
    channel group feats[:, channel_dims[i]:channel_dims[i+1]], all on the same rows (top-k-nearest points
    of each box, then superpoint means when spp_pool; the intersection is pooled too and the result
    broadcast back to its points, :243-249).  The two score columns are accumulated exactly as there
    (:236-237: column 1 receives max(p, 1-p) and column 0 min(p, 1-p) of every member, so the returned label
    is 1 unless the columns tie) - the mirror keeps that behaviour.  Returns (pred_probs, pred_labels,
    pred_variance) with pred_variance the summed Bernoulli variances.  `init_noise`: one array per group."""
    if not feats.is_cuda:
        raise _lib.GaproError("fit_gp_ensemble needs CUDA tensors; gapro_b200 has no CPU fallback")
    feats = feats.float()
    centroid = coords_float[intersect_inds].mean(0)
    k1 = _nearest_rows(coords_float[b1_inds], centroid, npoint_nearest)
    k2 = _nearest_rows(coords_float[b2_inds], centroid, npoint_nearest)
    b1_feats, b2_feats, test = feats[b1_inds][k1], feats[b2_inds][k2], feats[intersect_inds]
    inverse = None
    if spp_pool:
        b1_feats = _pool_by_superpoint(b1_feats, spp[b1_inds][k1])
        b2_feats = _pool_by_superpoint(b2_feats, spp[b2_inds][k2])
        test, inverse = _pool_by_superpoint(test, spp[intersect_inds], return_inverse=True)
    n1, n2, nt = int(b1_feats.shape[0]), int(b2_feats.shape[0]), int(test.shape[0])
    table = torch.cat([b1_feats, b2_feats, test], 0)
    score = torch.zeros((nt, 2), dtype=torch.float32, device=feats.device)
    variance = torch.zeros(nt, dtype=torch.float32, device=feats.device)
    for g in range(len(channel_dims) - 1):
        sub = table[:, channel_dims[g]:channel_dims[g + 1]].contiguous()
        probs, conf = fit_gp_regions(sub, [np.arange(n1 + n2)], [n1], [np.arange(n1 + n2, n1 + n2 + nt)],
                                     init_noise=None if init_noise is None else [init_noise[g]],
                                     training_iter=training_iter)[0][:2]
        score[:, 1] += conf
        score[:, 0] += 1 - conf
        variance += probs * (1 - probs)
    pred_probs, pred_labels = torch.max(score, dim=1)
    if inverse is not None:
        pred_probs, pred_labels, variance = pred_probs[inverse], pred_labels[inverse], variance[inverse]
    return pred_probs, pred_labels, variance
