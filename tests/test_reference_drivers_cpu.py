"""The reference's own driver scripts against the mirror API, without a GPU (build container only: needs /root/reference).

Every case*-script.py is passed through run_case.to_py3 (the mechanical Python-2 -> 3 rewrite), must compile, and every
name it takes from `from utils import *` / `from samplers import *` -- and every keyword it passes to HMC_sampler, gen_sample,
plot_samples and make_movie -- must exist in the mirror with that signature.  (Running them end to end needs a GPU;
tests/test_driver_gpu.py does that with a script of the same shape, since /root/reference is absent on the GPU box.)"""
import ast
import builtins
import glob
import inspect
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = json.load(open(os.path.join(ROOT, "BASELINE.json"))).get("reference_path", "/root/reference")
SCRIPTS = sorted(glob.glob(os.path.join(REF, "case*-script*.py")))

pytestmark = pytest.mark.skipif(not SCRIPTS, reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("path", SCRIPTS, ids=[os.path.basename(p) for p in SCRIPTS])
def test_reference_driver_resolves_against_mirror(path):
    sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
    import run_case
    import samplers as S
    import utils as U
    src = run_case.to_py3(open(path).read())
    tree = ast.parse(src, path)                                   # compiles under Python 3 after the mechanical rewrite
    exported = {n for n in dir(U) if not n.startswith("_")} | {n for n in dir(S) if not n.startswith("_")}
    assigned, used = set(), set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            (assigned if isinstance(node.ctx, (ast.Store, ast.Del)) else used).add(node.id)
        elif isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            assigned.add(node.name)
            if isinstance(node, ast.FunctionDef):
                assigned.update(a.arg for a in node.args.args)
        elif isinstance(node, ast.arg):
            assigned.add(node.arg)
    missing = {n for n in used if n not in assigned and not hasattr(builtins, n) and n not in exported}
    assert not missing, "names the driver expects from utils / samplers: %s" % sorted(missing)
    sigs = {"HMC_sampler": inspect.signature(S.HMC_sampler.__init__), "gen_sample": inspect.signature(S.HMC_sampler.gen_sample),
            "plot_samples": inspect.signature(S.HMC_sampler.plot_samples), "make_movie": inspect.signature(S.HMC_sampler.make_movie),
            "start_pts": inspect.signature(U.start_pts), "compute_convergence_stats": inspect.signature(S.HMC_sampler.compute_convergence_stats)}
    ncalls = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Call):
            name = node.func.id if isinstance(node.func, ast.Name) else (node.func.attr if isinstance(node.func, ast.Attribute) else None)
            if name in sigs:
                ncalls += 1
                params = sigs[name].parameters
                for kw in node.keywords:
                    assert kw.arg in params, "%s(%s=...) at line %d is not accepted by the mirror" % (name, kw.arg, node.lineno)
                npos = len(node.args) + (0 if name in ("HMC_sampler", "start_pts") and isinstance(node.func, ast.Name) else 1)
                assert npos <= len([p for p in params.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]) - \
                    (1 if name == "HMC_sampler" else 0) + (1 if name == "HMC_sampler" else 0)
    assert ncalls >= 3
