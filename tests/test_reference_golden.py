"""Golden vectors from the UNMODIFIED Julia reference (tests/golden/dump_reference.jl).

Julia cannot run in the build container or on the GPU box, so the directory tests/golden/reference/ does
not exist yet and these tests SKIP; the moment a maintainer runs the dump script and commits its output,
the CPU oracle (not gpu) and the CUDA path (gpu) are both held to the reference's own numbers: fixed-step
cases at 1e-5 (step-for-step), adaptive cases at solver tolerance, gradients against the reference's
tolerance-1e-4 continuous adjoint at 2e-3 (tests/test_oracle_adjoint.py measures that gap)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference")
MANIFEST = os.path.join(REF, "manifest.json")
have_ref = os.path.exists(MANIFEST)
needs_ref = pytest.mark.skipif(not have_ref, reason="tests/golden/reference/ absent: run tests/golden/dump_reference.jl with Julia")


def load_cases(manifest=MANIFEST, root=REF):
    cases = json.load(open(manifest))["cases"]
    for c in cases:
        D, B = c["nvariables"] + c["naugments"], c["B"]
        rd = lambda f, shape: np.fromfile(os.path.join(root, f"{c['name']}_{f}.f32"), dtype="<f4").reshape(shape, order="F")
        c["arrays"] = {"theta": rd("theta", (-1,)), "xs": rd("xs", (c["nvariables"], B)), "eps": rd("eps", (D, B)),
                       "logp_test": rd("logp_test", (B,)), "logp_train": rd("logp_train", (B,)), "E": rd("E", (B,)),
                       "n": rd("n", (B,)), "A": rd("A", (B,)), "loss": rd("loss", (1,)), "dtheta": rd("dtheta", (-1,)),
                       "dxs": rd("dxs", (c["nvariables"], B)), "z0": rd("z0", (D, B)), "generate": rd("generate", (c["nvariables"], B))}
        if c["nconditions"]:
            c["arrays"]["ys"] = rd("ys", (c["nconditions"], B))
    return cases


def tolerances(c):
    return (1e-5, 2e-3) if not c["adaptive"] else (1e-3, 5e-3)      # (values, gradients vs the continuous adjoint)


def _oracle(c):
    lam = c["lambda"]
    return O.OracleICNF(nvars=c["nvariables"], naug=c["naugments"], ncond=c["nconditions"], hidden=tuple(c["sizes"][1:-1]),
                        lam1=lam[0], lam2=lam[1], lam3=lam[2], steer_rate=c["steer_rate"])


def _opts(c):
    return O.SolverOpts(adaptive=bool(c["adaptive"]), dt=float(c["dt"]), reltol=c["reltol"], abstol=c["abstol"])


def _nrm(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(b) + 1e-30))


def test_loader_reads_a_dump_written_in_the_scripts_format(tmp_path):
    """always runs: the loader and the file format of dump_reference.jl agree (a synthetic dump written here)"""
    B, nv, na = 5, 2, 1
    names = {"theta": 7, "xs": nv * B, "eps": (nv + na) * B, "logp_test": B, "logp_train": B, "E": B, "n": B, "A": B, "loss": 1,
             "dtheta": 7, "dxs": nv * B, "z0": (nv + na) * B, "generate": nv * B}
    for f, n in names.items():
        np.arange(n, dtype="<f4").tofile(tmp_path / f"t_{f}.f32")
    man = {"cases": [{"name": "t", "nvariables": nv, "naugments": na, "nconditions": 0, "sizes": [4, 1, 3], "B": B, "seed": 1,
                      "adaptive": False, "dt": 0.5, "reltol": 1e-4, "abstol": 1e-4, "steer_rate": 0.0, "t0": 0.0, "t1": 1.0,
                      "lambda": [0.01, 0.01, 0.01]}]}
    (tmp_path / "manifest.json").write_text(json.dumps(man))
    c = load_cases(str(tmp_path / "manifest.json"), str(tmp_path))[0]
    assert c["arrays"]["xs"].shape == (nv, B) and c["arrays"]["xs"][1, 0] == 1.0 and c["arrays"]["xs"][0, 1] == 2.0   # column-major


@needs_ref
def test_oracle_matches_the_reference():
    for c in load_cases():
        om, a, vt, gt = _oracle(c), c["arrays"], *tolerances(c)
        t = lambda x: None if x is None else torch.tensor(np.asarray(x), dtype=torch.float64)
        ys = t(a.get("ys"))
        lp, _ = O.inference(om, O.TEST, t(a["xs"]), t(a["theta"]), None, ys, opts=_opts(c))
        assert _nrm(lp.numpy(), a["logp_test"]) < vt, (c["name"], "logp_test")
        lp, (E, n, A) = O.inference(om, O.TRAIN_REG, t(a["xs"]), t(a["theta"]), t(a["eps"]), ys, t1=c["t1"], opts=_opts(c))
        assert _nrm(lp.numpy(), a["logp_train"]) < vt, (c["name"], "logp_train")
        assert _nrm(E.numpy(), a["E"]) < 10 * vt and _nrm(n.numpy(), a["n"]) < 10 * vt
        l, g, gx = O.loss_grad(om, O.TRAIN_REG, t(a["xs"]), t(a["theta"]), t(a["eps"]), ys, t1=c["t1"], opts=_opts(c), want_dxs=True)
        assert abs(float(l) - a["loss"][0]) < vt * abs(a["loss"][0]) + 1e-6
        assert _nrm(g.numpy(), a["dtheta"]) < gt and _nrm(gx.numpy(), a["dxs"]) < gt, c["name"]
        gen = O.generate(om, O.TEST, t(a["z0"]), t(a["theta"]), None, ys, opts=_opts(c))
        assert _nrm(gen.numpy(), a["generate"]) < vt, (c["name"], "generate")


@needs_ref
@pytest.mark.gpu
def test_cuda_path_matches_the_reference():
    import cnf_b200 as m
    for c in load_cases():
        a, (vt, gt) = c["arrays"], tolerances(c)
        lam = c["lambda"]
        icnf = m.ICNF(nvariables=c["nvariables"], naugments=c["naugments"], nconditions=c["nconditions"],
                      n_hidden=c["sizes"][1], lambda1=lam[0], lambda2=lam[1], lambda3=lam[2], steer_rate=c["steer_rate"])
        sol = dict(adaptive=bool(c["adaptive"]), dt=float(c["dt"]), reltol=c["reltol"], abstol=c["abstol"])
        ya = (a["ys"],) if c["nconditions"] else ()
        lp, _ = m.inference(icnf, m.TestMode(), a["xs"], *ya, a["theta"], {}, tspan=(c["t0"], 1.0), **sol)
        assert _nrm(lp, a["logp_test"]) < max(vt, 1e-4), (c["name"], "logp_test")
        lp, (E, n, A) = m.inference(icnf, m.TrainMode(True), a["xs"], *ya, a["theta"], {}, eps=a["eps"], tspan=(c["t0"], c["t1"]), **sol)
        assert _nrm(lp, a["logp_train"]) < max(vt, 1e-4), (c["name"], "logp_train")
        l, g, gx = m.loss_and_gradient(icnf, m.TrainMode(True), a["xs"], *ya, a["theta"], {}, want_dxs=True, eps=a["eps"],
                                       tspan=(c["t0"], c["t1"]), **sol)
        assert abs(l - a["loss"][0]) < max(vt, 1e-4) * abs(a["loss"][0]) + 1e-5
        assert _nrm(g, a["dtheta"]) < gt and _nrm(gx, a["dxs"]) < gt, c["name"]
        gen = m.generate(icnf, m.TestMode(), *ya, a["theta"], {}, c["B"], z0=a["z0"], tspan=(c["t0"], 1.0), **sol)
        assert _nrm(gen, a["generate"]) < max(vt, 1e-4), (c["name"], "generate")
