"""TEST INFRASTRUCTURE ONLY -- executes the UNMODIFIED-IN-SUBSTANCE reference through a mechanical py2->py3 shim.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``, ``tests/golden/make_golden.py``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it.

The reference (``/root/reference``; jaekor91/understanding-HMC) is Python 2 and is mounted read-only; it is
*not* copied into this repo.  This module reads ``samplers.py`` / ``utils.py`` from ``reference_path`` at call
time, applies only the following text rewrites (no arithmetic is touched), and ``exec``s the
result into fresh module objects:

1. ``print <expr>`` statements -> ``print(<expr>)``        (samplers.py:424..870, utils.py:404)
2. ``xrange`` -> ``range``                                   (samplers.py:410, 428, 448, 545, 563, 637; utils.py:89)
3. ``np.float`` -> ``float``, ``np.int`` -> ``int``          (samplers.py:33, 359, 360, 399)
4. ``n = L_chain/2`` -> ``n = L_chain//2``                   (utils.py:102, python-2 integer division)
5. the invalid escape ``\\%d`` in a format string            (samplers.py:285, plotting only)
6. stub modules for matplotlib (absent in this image)       (utils.py:2, 12, 19)

The reference path does not exist on the GPU box, so everything here is optional at import time:
``available()`` says whether the reference can be loaded.
"""
from __future__ import annotations

import json
import os
import re
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def reference_path() -> str:
    """`BASELINE.json: reference_path` (falls back to /root/reference)."""
    try:
        with open(os.path.join(_ROOT, "BASELINE.json")) as f:
            return json.load(f).get("reference_path", "/root/reference")
    except Exception:
        return "/root/reference"


def available() -> bool:
    p = reference_path()
    return os.path.isfile(os.path.join(p, "samplers.py")) and os.path.isfile(os.path.join(p, "utils.py"))


_PRINT_RE = re.compile(r"^(\s*)print\s+(?!\()(.+?)\s*$")
_PRINT_PAREN_RE = re.compile(r"^(\s*)print\s+(\(.*\)\s*%.*)$")


def py2_to_py3(src: str) -> str:
    """The mechanical rewrite list of the module docstring, applied line by line."""
    out = []
    for line in src.split("\n"):
        stripped = line.lstrip()
        if not stripped.startswith("#"):
            m = _PRINT_RE.match(line) or _PRINT_PAREN_RE.match(line)
            if m:
                line = "%sprint(%s)" % (m.group(1), m.group(2))
            elif stripped.rstrip() == "print":
                line = line.replace("print", "print()")
        line = line.replace("xrange", "range")
        line = re.sub(r"\bnp\.float\b(?!\d|_)", "float", line)
        line = re.sub(r"\bnp\.int\b(?!\d|_|e)", "int", line)
        line = re.sub(r"\bn = L_chain/2\b", "n = L_chain//2", line)
        line = line.replace("\\%d", "/%d")
        out.append(line)
    return "\n".join(out)


def _stub_matplotlib():
    """Register inert matplotlib stubs if the real package is missing (plots are out of scope)."""
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.rcParams = {}
    plt = types.ModuleType("matplotlib.pyplot")
    patches = types.ModuleType("matplotlib.patches")
    patches.Ellipse = object
    mpl.pyplot = plt
    mpl.patches = patches
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("matplotlib.patches", patches)


_CACHE = {}


def load_reference(quiet: bool = True):
    """Return ``(ref_utils, ref_samplers)`` module objects built from the reference source text.

    ``quiet`` replaces the module-level ``print`` by a no-op (the reference prints once per chain).
    """
    key = (reference_path(), quiet)
    if key in _CACHE:
        return _CACHE[key]
    if not available():
        raise RuntimeError("reference not present at %s (expected on the build container only)" % reference_path())
    _stub_matplotlib()
    rp = reference_path()
    with open(os.path.join(rp, "utils.py")) as f:
        usrc = py2_to_py3(f.read())
    with open(os.path.join(rp, "samplers.py")) as f:
        ssrc = py2_to_py3(f.read())
    ref_utils = types.ModuleType("_ref_utils")
    ref_utils.__file__ = os.path.join(rp, "utils.py")
    if quiet:
        ref_utils.__dict__["print"] = lambda *a, **k: None
    exec(compile(usrc, ref_utils.__file__, "exec"), ref_utils.__dict__)
    ref_samplers = types.ModuleType("_ref_samplers")
    ref_samplers.__file__ = os.path.join(rp, "samplers.py")
    if quiet:
        ref_samplers.__dict__["print"] = lambda *a, **k: None
    # `from utils import *` at samplers.py:1 must resolve to the shimmed utils.
    saved = sys.modules.get("utils")
    sys.modules["utils"] = ref_utils
    try:
        exec(compile(ssrc, ref_samplers.__file__, "exec"), ref_samplers.__dict__)
    finally:
        if saved is None:
            sys.modules.pop("utils", None)
        else:
            sys.modules["utils"] = saved
    _CACHE[key] = (ref_utils, ref_samplers)
    return _CACHE[key]


class DrawRecorder(object):
    """Records the reference's draws in consumption order.

    Wraps ``sampler.p_sample`` plus ``np.random.randint`` / ``np.random.random`` *as seen by the reference
    module* so the tape can be replayed into the oracle restatement and into the CUDA kernels.
    Tape entries: ("p", array(D)), ("i", int), ("u", float).
    """

    def __init__(self, ref_samplers_mod, sampler_obj):
        self.tape = []
        self._mod = ref_samplers_mod
        self._obj = sampler_obj
        self._np = ref_samplers_mod.np

    def __enter__(self):
        import numpy as real_np
        rec = self

        orig_p = self._obj.p_sample

        def p_sample():
            out = orig_p()
            rec.tape.append(("p", real_np.array(out[0], dtype=float)))
            return out

        self._obj.p_sample = p_sample

        class _RandomProxy(object):
            def __getattr__(self, name):
                return getattr(real_np.random, name)

            def randint(self, *a, **k):
                out = real_np.random.randint(*a, **k)
                rec.tape.append(("i", int(real_np.asarray(out).ravel()[0])))
                return out

            def random(self, *a, **k):
                out = real_np.random.random(*a, **k)
                rec.tape.append(("u", float(real_np.asarray(out).ravel()[0])))
                return out

        class _NpProxy(object):
            random = _RandomProxy()

            def __getattr__(self, name):
                return getattr(real_np, name)

        self._saved_np = self._mod.__dict__["np"]
        self._mod.__dict__["np"] = _NpProxy()
        return self

    def __exit__(self, *exc):
        self._mod.__dict__["np"] = self._saved_np
        try:
            del self._obj.p_sample
        except Exception:
            pass
        return False
