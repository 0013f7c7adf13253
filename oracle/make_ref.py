#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY -- makes the reference itself runnable on the GPU box.

The reference (walker-gym) is pure Python with no build system and no package metadata, so there is nothing to
``pip install``.  This recipe byte-compiles the four UNMODIFIED modules of its "optimized flat" lineage

    gym/optimized_engine.py  gym/optimized_renderer.py  gym/optimized_walker.py  gym/optimized_env.py

from the read-only checkout (default /root/reference) into ``oracle/_ref/gym/*.refbc`` (CPython bytecode files under a neutral extension, so that
snapshot tools that drop ``*.pyc`` still ship them) -- bytecode only, the Python
analogue of compiling a C reference into ``oracle/_ref/*.so``: no reference source text enters the repository, the
directory is git-ignored, and it travels to the GPU box with the snapshot (same image, same interpreter, same
bytecode magic).  ``oracle/ref_harness.py`` imports the modules from there (pygame / turtle stubbed, the six-line
``Point.forced`` shim -- SURVEY.md section 0.3), and ``bench.py --impl reference`` times them on the box's host cores.

    python oracle/make_ref.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "gym")
MODULES = ("optimized_engine", "optimized_renderer", "optimized_walker", "optimized_env")


def make(ref_root: str = "/root/reference", quiet: bool = False) -> bool:
    src_dir = os.path.join(ref_root, "gym")
    if not all(os.path.isfile(os.path.join(src_dir, m + ".py")) for m in MODULES):
        if not quiet:
            print(f"make_ref: no reference checkout under {ref_root}; keeping whatever {OUT} holds")
        return False
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(src_dir, m + ".py"), cfile=os.path.join(OUT, m + ".refbc"),
                           dfile=f"<reference>/gym/{m}.py", doraise=True, optimize=0)
    with open(os.path.join(OUT, "PROVENANCE"), "w") as f:
        f.write(f"bytecode of {', '.join(m + '.py' for m in MODULES)} from {src_dir}, unmodified; "
                f"python {sys.version.split()[0]}; made by oracle/make_ref.py\n")
    if not quiet:
        print(f"make_ref: {len(MODULES)} modules -> {OUT}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("WALKER_GYM_REFERENCE", "/root/reference"))
    make(ap.parse_args().ref)
