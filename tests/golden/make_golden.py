"""Regenerates the committed golden vectors.  The reference itself cannot be imported here
(gpytorch / torch_scatter are absent), so these vectors are OUTPUTS OF THE ORACLE
(fp64 policy) — they pin the oracle against accidental drift and let the GPU box check the CUDA
path without re-running the slow oracle; they do NOT pin the oracle to the reference (that is
make_ref_golden.py's job for the scene pipeline; the GP fit stays "parity unpinned", see oracle/__init__.py).

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from gapro_b200 import synthetic                       # noqa: E402
from gapro_b200.gen_ps import synthetic_inputs         # noqa: E402
from oracle import gen_ps_oracle as O                  # noqa: E402
from oracle import gp_oracle as G                      # noqa: E402

GP_CASES = [(2, 6, 1), (3, 6, 4), (17, 6, 9), (64, 6, 30), (65, 6, 70), (96, 32, 12), (150, 6, 20)]
# sizes of the real workload (16 blocks of 64, deep features): golden vectors only, the CPU suite re-derives the first
GP_CASES_LARGE = [(200, 6, 40), (1000, 6, 64), (520, 32, 48)]


def gp_case(i, M, D, N):
    rng = np.random.default_rng([42, i])
    n1 = max(1, M // 3)
    c1 = rng.normal(size=D)
    c2 = c1 + rng.normal(size=D) * 0.8
    X = np.concatenate([c1 + 0.5 * rng.normal(size=(n1, D)), c2 + 0.5 * rng.normal(size=(M - n1, D))]).astype(np.float32)
    Xt = (0.5 * (c1 + c2) + 0.5 * rng.normal(size=(N, D))).astype(np.float32)
    noise = rng.standard_normal(M).astype(np.float32)
    return X, n1, Xt, noise


def scene_inputs(name, seed):
    inp = synthetic_inputs(synthetic.make_scene(seed, name))
    return inp, (inp["xyz"], inp["mask_feats"].astype(np.float32), inp["spp"], inp["instance_cls"].astype(np.int64),
                 inp["instance_box"].astype(np.float32), inp["instance_box_volume"].astype(np.float32),
                 inp["wall_box"], inp["wall_volume"])


def input_digest(args):
    h = hashlib.sha256()
    for a in args:
        h.update(np.ascontiguousarray(np.asarray(a)).tobytes())
    return h.hexdigest()


def main():
    out = {}
    for i, (M, D, N) in enumerate(GP_CASES):
        X, n1, Xt, noise = gp_case(i, M, D, N)
        r = G.fit_region_autograd(X, n1, Xt, noise)
        out[f"c{i}_mu64"], out[f"c{i}_var64"], out[f"c{i}_prob"] = r["mu64"], r["var64"], r["prob"]
        out[f"c{i}_label"], out[f"c{i}_conf"] = r["label"], r["conf"]
        print(f"gp case {i}: M={M} D={D} N={N} mu[0]={r['mu64'][0]:.6f}")
    np.savez_compressed(os.path.join(HERE, "gp_cases.npz"), **out)
    out = {}
    for i, (M, D, N) in enumerate(GP_CASES_LARGE):
        X, n1, Xt, noise = gp_case(100 + i, M, D, N)
        r = G.fit_region_autograd(X, n1, Xt, noise)
        out[f"c{i}_mu64"], out[f"c{i}_var64"], out[f"c{i}_prob"] = r["mu64"], r["var64"], r["prob"]
        out[f"c{i}_label"], out[f"c{i}_conf"] = r["label"], r["conf"]
        print(f"large gp case {i}: M={M} D={D} N={N} mu[0]={r['mu64'][0]:.6f} min margin {np.abs(r['prob64'] - 0.5).min():.3g}")
    np.savez_compressed(os.path.join(HERE, "gp_cases_large.npz"), **out)
    for name, seed, nseed in [("tiny", 3, 5), ("small", 4, 6)]:
        inp, args = scene_inputs(name, seed)
        res, dbg = O.gen_pseudo_label_oracle(*args, thresh_spp_occu=0.999, noise_seed=nseed, return_debug=True)
        margins = []
        for r in dbg["regions"]:
            margins.append(np.abs(r["res"]["prob64"] - 0.5).min())
        np.savez_compressed(os.path.join(HERE, f"scene_{name}.npz"), sem=res[0], inst=res[1], prob=res[2], mu=res[3],
                            var=res[4], digest=np.array(input_digest(args)), n_regions=np.array(len(dbg["regions"])),
                            min_margin=np.array(min(margins) if margins else 1.0),
                            occ_spp=np.packbits(dbg["occ_spp"], axis=1), feats_spp=dbg["feats_spp"])
        print(f"scene {name}: {len(res[0])} pts, {len(res[3])} spp, {len(dbg['regions'])} GP regions, "
              f"{sum(e[0] == 'nest' for e in dbg['events'])} nest events, min posterior margin {min(margins) if margins else 1:.3g}")


if __name__ == "__main__":
    main()
