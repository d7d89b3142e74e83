"""Test hook around gapro_gp_debug_run: run one region for `iters` steps plus the first
`stop_phase` phases of the next step and read its workspace buffers back as numpy arrays."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

PHASES = ["build", "chol", "A", "B", "colstats", "GA", "GT", "GM", "GC", "GL", "SP", "Y", "GK", "kgrad", "adam"]


def gp_debug_state(feats_spp, train_idx, n_b1, test_idx, init_noise, iters=0, stop_phase=0, lr=0.1,
                   jitter_zz=1e-4, jitter_xx=1e-4):
    lib = _lib.load()
    dev = feats_spp.device
    feats = feats_spp.float().contiguous()
    D = int(feats.shape[1])
    tr = torch.as_tensor(np.asarray(train_idx, dtype=np.int32)).to(dev)
    te = torch.as_tensor(np.asarray(test_idx, dtype=np.int32)).to(dev)
    nz = torch.as_tensor(np.asarray(init_noise, dtype=np.float32)).to(dev)
    M, N = int(tr.numel()), int(te.numel())
    off = np.array([0, M], dtype=np.int32)
    toff = np.array([0, N], dtype=np.int32)
    nbytes = lib.gapro_gp_workspace_bytes(1, off.ctypes.data, toff.ctypes.data, D) + 4096
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    layout = np.zeros(32, dtype=np.int64)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.gapro_gp_debug_run(feats.data_ptr(), D, M, int(n_b1), N, tr.data_ptr(), te.data_ptr(), nz.data_ptr(),
                                      int(iters), int(stop_phase), float(lr), float(jitter_zz), float(jitter_xx),
                                      ws.data_ptr(), ws.numel(), layout.ctypes.data, 32, stream), "gapro_gp_debug_run")
    torch.cuda.synchronize(dev)
    names = lib.gapro_gp_debug_layout_names().decode().split(",")
    lay = dict(zip(names, layout.tolist()))
    Mp, Np, Wp = lay["Mp"], lay["Np"], lay["Wp"]
    buf = ws[: lay["total"] * 8].view(torch.float64).cpu().numpy()
    status = int(ws[nbytes - 64: nbytes - 60].view(torch.int32).item())

    def mat(name, rows, ld, r=None, c=None):
        a = buf[lay[name]: lay[name] + rows * ld].reshape(rows, ld)
        return a[: (r or rows), : (c or ld)].copy()

    out = dict(status=status, Mp=Mp, Np=Np, Wp=Wp)
    for nm in ("X", "Z", "Zm", "Zv", "gZ"):
        out[nm] = mat(nm, Mp, D, M, D)
    for nm in ("y", "m", "mm", "mv", "gsrow", "glrow"):
        out[nm] = buf[lay[nm]: lay[nm] + M].copy()
    for nm in ("mu", "var", "gmu", "gv"):
        out[nm] = buf[lay[nm]: lay[nm] + Wp].copy()
    out["scal"] = buf[lay["scal"]: lay["scal"] + 16].copy()
    for nm in ("L", "Linv", "T", "Tm", "Tv", "GA", "GC"):
        out[nm] = mat(nm, Mp, Mp)
    for nm in ("Kzx", "A", "Bm"):
        out[nm] = mat(nm, Mp, Wp)
    return out
