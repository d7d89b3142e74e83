"""Stage-by-stage GPU diagnostic (development aid, run under gpurun):
    python tests/gpu_diag.py [section ...]  > gpurun_out/diag.txt
Every section is independent and failures are reported, not fatal."""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gapro_b200 import _debug, _lib, synthetic                      # noqa: E402
from gapro_b200.engine import get_engine                            # noqa: E402
from gapro_b200.gaussian_process_utils import fit_gp_regions        # noqa: E402
from gapro_b200.gen_ps import synthetic_inputs, to_scene_inputs     # noqa: E402
from oracle import gen_ps_oracle as O                               # noqa: E402
from oracle import gp_oracle as G                                   # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def oracle_args(inp):
    return (inp["xyz"], inp["mask_feats"].astype(np.float32), inp["spp"], inp["instance_cls"].astype(np.int64),
            inp["instance_box"].astype(np.float32), inp["instance_box_volume"].astype(np.float32),
            inp["wall_box"], inp["wall_volume"])


def unpack_bits(bits, B):
    bits = bits.cpu().numpy().view(np.uint32)
    S, W = bits.shape
    out = np.zeros((S, 32 * W), dtype=bool)
    for w in range(W):
        for b in range(32):
            out[:, 32 * w + b] = (bits[:, w] >> b) & 1
    return out[:, :B]


def section_stages():
    print("== stages: tiny + small as one batch, against the oracle")
    eng = get_engine(DEV)
    names = ["tiny", "small"]
    inps = [synthetic_inputs(synthetic.make_scene(7 + i, n)) for i, n in enumerate(names)]
    scenes = [to_scene_inputs(inp, DEV, noise_seed=11 + i) for i, inp in enumerate(inps)]
    fake = lambda X, n1, Xt, nz: dict(conf=np.full(len(Xt), 0.75, np.float32), label=np.ones(len(Xt), bool),
                                      mu=np.zeros(len(Xt), np.float32), var=np.ones(len(Xt), np.float32))
    outs, dbg = eng.run(scenes, thresh_spp_occu=0.999, training_iter=0, debug=True, want_cnt_in=True)
    torch.cuda.synchronize()
    for i, inp in enumerate(inps):
        _, od = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, fit_fn=fake, return_debug=True)
        p0 = sum(len(x["xyz"]) for x in inps[:i])
        s0, s1 = dbg.spp_off[i], dbg.spp_off[i + 1]
        b0, b1 = dbg.box_off[i], dbg.box_off[i + 1]
        B = b1 - b0
        gid = dbg.spp_gid.cpu().numpy()[p0:p0 + len(inp["xyz"])] - s0
        print(f" scene {names[i]}: S {s1 - s0} vs {len(od['n_bbs'])}; dense ids equal:",
              bool((gid == od['spp_dense']).all()))
        print("  boxes equal:", bool((dbg.boxes[b0:b1] == od['boxes']).all()), " vol equal:",
              bool((dbg.boxes_vol[b0:b1] == od['boxes_vol']).all()))
        ci = dbg.cnt_in.cpu().numpy()[s0:s1, :B]
        print("  cnt_in equal:", bool((ci == od['cnt_in']).all()))
        occ = unpack_bits(dbg.occ_bits[s0:s1], B)
        print("  occ_spp equal:", bool((occ == od['occ_spp']).all()), " n_bbs equal:",
              bool((dbg.n_bbs.cpu().numpy()[s0:s1] == od['n_bbs']).all()))
        f = dbg.feats_spp.cpu().numpy()[s0:s1]
        print("  feats_spp bit-exact:", bool((f.view(np.uint32) == od['feats_spp'].view(np.uint32)).all()),
              " max abs diff", float(np.abs(f - od['feats_spp']).max()))
        kinds = {0: "nest", 1: "nest", 2: "gp"}
        ev = [(kinds[k], a, b) for k, a, b in dbg.events[i]]
        oev = [(e[0], e[1], e[2]) for e in od['events']]
        print("  events equal:", ev == oev, f"({len(ev)} events, {sum(e[0] == 'gp' for e in ev)} gp)")
        regs = [r for r in dbg.regions if r["scene"] == i]
        ok = all((r["train_idx"] == np.concatenate([o["b1_inds"], o["b2_inds"]])).all() and
                 (r["test_idx"] == o["inter"]).all() and (r["noise"] == o["noise"]).all()
                 for r, o in zip(regs, od["regions"])) if len(regs) == len(od["regions"]) else False
        print("  region index lists + noise equal:", ok)


def pick_region(name="small", seed=8, which="largest"):
    inp = synthetic_inputs(synthetic.make_scene(seed, name))
    fake = lambda X, n1, Xt, nz: dict(conf=np.full(len(Xt), 0.75, np.float32), label=np.ones(len(Xt), bool),
                                      mu=np.zeros(len(Xt), np.float32), var=np.ones(len(Xt), np.float32))
    _, od = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, fit_fn=fake, return_debug=True)
    regs = od["regions"]
    sizes = [len(r["b1_inds"]) + len(r["b2_inds"]) for r in regs]
    r = regs[int(np.argmax(sizes))] if which == "largest" else regs[int(np.argmin(sizes))]
    return od["feats_spp"], r


def section_phases():
    print("== GP phases of step 1 against the numpy mirror")
    feats, r = pick_region()
    train = np.concatenate([r["b1_inds"], r["b2_inds"]])
    n1, test = len(r["b1_inds"]), r["inter"]
    M, N = len(train), len(test)
    print(f" region M={M} n_b1={n1} N_int={N}")
    rng = np.random.default_rng(3)
    noise = rng.standard_normal(M).astype(np.float32)
    X = feats[train].astype(np.float64)
    y = np.concatenate([-np.ones(n1), np.ones(M - n1)])
    Z, m, T = X.copy(), 1e-3 * noise.astype(np.float64), np.eye(M)
    grads, it = G.manual_grads(Z, m, T, 0.0, 0.0, 0.0, X, y, 1e-4, 1e-4, return_internals=True)
    fd = torch.from_numpy(feats).to(DEV)
    st = lambda p: _debug.gp_debug_state(fd, train, n1, test, noise, iters=0, stop_phase=p)
    ph = {n: i for i, n in enumerate(_debug.PHASES)}
    s = st(ph["build"] + 1)
    ell = s_ = np.log(2.0)
    Kzz = G._rbf(Z, Z, ell, s_)[0] + 1e-4 * np.eye(M)
    print("  build Kzx:", rel(s["Kzx"][:M, :M], it["Kzx"]), " Kzz(lower):", rel(np.tril(s["L"][:M, :M]), np.tril(Kzz)),
          " pad diag ok:", bool(np.all(np.diag(s["L"])[M:] == 1.0)))
    s = st(ph["chol"] + 1)
    print("  chol L:", rel(np.tril(s["L"][:M, :M]), it["L"]), " Linv:", rel(s["Linv"][:M, :M], it["Linv"]),
          " status", s["status"], " upper zero:", float(np.abs(np.triu(s["Linv"], 1)).max()))
    s = st(ph["A"] + 1)
    print("  A:", rel(s["A"][:M, :M], it["A"]))
    s = st(ph["B"] + 1)
    print("  B:", rel(s["Bm"][:M, :M], it["B"]))
    s = st(ph["colstats"] + 1)
    print("  mu:", rel(s["mu"][:M], it["mu"]), " var:", rel(s["var"][:M], it["var"]), " g_mu:",
          rel(s["gmu"][:M], it["g_mu"]), " g_v:", rel(s["gv"][:M], it["g_v"]))
    s = st(ph["GA"] + 1)
    print("  G_A:", rel(s["GA"][:M, :M], it["G_A"]))
    # after GT / GM the parameters T and m took their Adam step
    opt = G._Adam([Z.copy(), m.copy(), T.copy(), np.zeros(()), np.zeros(()), np.zeros(())], 0.1)
    opt.step([np.asarray(g) for g in grads])
    Z1, m1, T1, c1, rs1, rl1 = opt.params
    s = st(ph["GT"] + 1)
    print("  T after Adam:", rel(s["T"][:M, :M], T1), " pad T identity:",
          bool(np.all(np.diag(s["T"])[M:] == 1.0)), " upper zero:", float(np.abs(np.triu(s["T"], 1)).max()))
    s = st(ph["GM"] + 1)
    print("  m after Adam:", rel(s["m"], m1))
    s = st(ph["GC"] + 1)
    print("  G_C:", rel(s["GC"][:M, :M], it["G_C"]))
    s = st(ph["GL"] + 1)
    print("  symP:", rel(s["Bm"][:M, :M], it["symP"]))
    s = st(ph["Y"] + 1)
    print("  Y (lower tiles):", rel(np.tril(s["GA"][:M, :M]), np.tril(it["symP"] @ it["Linv"])))
    s = st(ph["GK"] + 1)
    print("  G_K:", rel(s["Bm"][:M, :M], it["G_K"]))
    s = st(ph["kgrad"] + 1)
    print("  gZ:", rel(s["gZ"], grads[0]))
    s = st(ph["adam"] + 1)
    print("  after step 1: Z", rel(s["Z"], Z1), " c", abs(s["scal"][0] - c1), " rho_s", abs(s["scal"][1] - rs1),
          " rho_l", abs(s["scal"][2] - rl1))
    # three full steps
    for _ in range(2):
        g = G.manual_grads(Z1, m1, T1, float(c1), float(rs1), float(rl1), X, y, 1e-4, 1e-4)
        opt.step([np.asarray(x) for x in g])
    s = _debug.gp_debug_state(fd, train, n1, test, noise, iters=3, stop_phase=0)
    print("  after 3 steps: Z", rel(s["Z"], opt.params[0]), " m", rel(s["m"], opt.params[1]), " T",
          rel(s["T"][:M, :M], opt.params[2]), " scal", np.abs(s["scal"][:3] - np.array([float(p) for p in opt.params[3:]])).max())


def section_fits():
    print("== full 50-step fits against the fp64 autograd oracle (tolerance rel 1e-4)")
    rng = np.random.default_rng(5)
    cases = []
    for M, D, N in [(2, 6, 1), (3, 6, 4), (17, 6, 9), (64, 6, 30), (65, 6, 70), (200, 6, 33), (130, 32, 20)]:
        n1 = max(1, M // 3)
        c1 = rng.normal(size=D)
        c2 = c1 + rng.normal(size=D) * 0.8
        X = np.concatenate([c1 + 0.5 * rng.normal(size=(n1, D)), c2 + 0.5 * rng.normal(size=(M - n1, D))])
        Xt = 0.5 * (c1 + c2) + 0.5 * rng.normal(size=(N, D))
        cases.append((X.astype(np.float32), n1, Xt.astype(np.float32), rng.standard_normal(M).astype(np.float32)))
    # duplicated rows (exercise the jitter)
    X, n1, Xt, nz = cases[2]
    cases.append((np.concatenate([X, X[:5]]), n1, Xt, np.concatenate([nz, nz[:5]])))
    for D in (6, 32):
        sub = [c for c in cases if c[0].shape[1] == D]
        feats = np.concatenate([np.concatenate([c[0], c[2]]) for c in sub])
        tr, te, off = [], [], 0
        for c in sub:
            tr.append(np.arange(off, off + len(c[0])))
            te.append(np.arange(off + len(c[0]), off + len(c[0]) + len(c[2])))
            off += len(c[0]) + len(c[2])
        t0 = time.time()
        res = fit_gp_regions(torch.from_numpy(feats).to(DEV), tr, [c[1] for c in sub], te,
                             init_noise=[c[3] for c in sub], return_float64=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        for c, r in zip(sub, res):
            o = G.fit_region_autograd(c[0], c[1], c[2], c[3])
            mu, var = r[5].cpu().numpy(), r[6].cpu().numpy()
            print(f"  M={len(c[0])} D={D} N={len(c[2])}: mu rel {rel(mu, o['mu64']):.2e} var rel {rel(var, o['var64']):.2e}"
                  f" labels equal {bool((r[2].cpu().numpy() == o['label']).all())} conf max diff "
                  f"{float(np.abs(r[1].cpu().numpy() - o['conf']).max()):.2e}")
        print(f"  (batch of {len(sub)} regions, D={D}: {dt * 1e3:.1f} ms wall incl. setup)")


def section_e2e():
    print("== end-to-end scene parity (oracle fp64 GP)")
    eng = get_engine(DEV)
    for name, seed in [("tiny", 21), ("small", 22)]:
        inp = synthetic_inputs(synthetic.make_scene(seed, name))
        t0 = time.time()
        ref, od = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, noise_seed=99, return_debug=True)
        t_or = time.time() - t0
        sc = to_scene_inputs(inp, DEV, noise_seed=99)
        out = eng.run([sc], thresh_spp_occu=0.999)[0]
        torch.cuda.synchronize()
        sem, inst, prob, mu, var = [t.cpu().numpy() for t in out]
        ok = mu != -100
        print(f"  {name}: sem equal {bool((sem == ref[0]).all())} inst equal {bool((inst == ref[1]).all())} "
              f"prob max diff {float(np.abs(prob - ref[2]).max()):.2e} sentinel equal {bool(((mu == -100) == (ref[3] == -100)).all())} "
              f"mu rel {rel(mu[ok], ref[3][ok]) if ok.any() else 0:.2e} var rel {rel(var[ok], ref[4][ok]) if ok.any() else 0:.2e} "
              f"(oracle {t_or:.1f}s, {len(od['regions'])} regions)")


def section_time():
    print("== c1 scene timing (1 scene, then 4 scenes)")
    eng = get_engine(DEV)
    inps = [synthetic_inputs(synthetic.make_scene(100 + i, "c1")) for i in range(4)]
    scenes = [to_scene_inputs(inp, DEV, noise_seed=i) for i, inp in enumerate(inps)]
    for batch in (scenes[:1], scenes[:1], scenes):
        torch.cuda.synchronize()
        t0 = time.time()
        eng.run(batch, thresh_spp_occu=0.999)
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = eng.last_stats
        print(f"  {len(batch)} scene(s): {dt * 1e3:.1f} ms; regions {st['n_regions']} sumM {st['sum_m']} "
              f"sumM3 {st['sum_m3']:.3g} launches {st['launches']} -> {433 * st['sum_m3'] / dt / 1e12:.2f} TFLOP/s fp64 (algorithmic)")


SECTIONS = dict(stages=section_stages, phases=section_phases, fits=section_fits, e2e=section_e2e, time=section_time)

if __name__ == "__main__":
    todo = sys.argv[1:] or list(SECTIONS)
    print("device:", torch.cuda.get_device_name(0), "| lib version", _lib.load().gapro_version())
    for name in todo:
        t0 = time.time()
        try:
            SECTIONS[name]()
        except Exception:
            print(f"!! section {name} FAILED")
            traceback.print_exc(file=sys.stdout)
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("!! CUDA error after section", name, e)
            break
        print(f"   [{name}: {time.time() - t0:.1f}s]", flush=True)
