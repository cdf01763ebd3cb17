#!/usr/bin/env python
"""Runs one of the reference's Python-2 driver scripts (case*-script.py) unchanged against the B200 build.

    python understanding-hmc_b200/run_case.py /path/to/case1-script.py [--dtype float64] [--max-cases N]

The script text is read, the Python-2-only syntax is rewritten mechanically (`print x` statements, `xrange`), and
the result is exec'ed with this directory first on sys.path so that `from utils import *` / `from samplers import *`
(case1-script.py:1-2) resolve to the CUDA-backed mirrors.  Nothing else is touched: the script's own constants,
target closures, `start_pts` calls and prints run as written."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_PRINT = re.compile(r"^(\s*)print\s+(?!\()(.+?)\s*$")
_PRINT_FMT = re.compile(r"^(\s*)print\s+(\(.*\)\s*%.*)$")


def to_py3(src):
    out = []
    for line in src.split("\n"):
        if not line.lstrip().startswith("#"):
            m = _PRINT.match(line) or _PRINT_FMT.match(line)
            if m:
                line = "%sprint(%s)" % (m.group(1), m.group(2))
            elif line.strip() == "print":
                line = line.replace("print", "print()")
        out.append(line.replace("xrange", "range"))
    return "\n".join(out)


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    path = argv[1]
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    with open(path) as f:
        src = to_py3(f.read())
    os.makedirs(os.path.join(os.getcwd(), os.path.splitext(os.path.basename(path))[0].split("-")[0]), exist_ok=True)
    glb = {"__name__": "__main__", "__file__": path}
    exec(compile(src, path, "exec"), glb)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
