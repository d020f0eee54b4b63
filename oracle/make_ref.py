#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Materialise a runnable copy of the (Python 2.7) reference under
``oracle/_ref/BNMTF_ARD/`` so that it can be imported by
  * ``oracle/gen_golden.py``   (golden fixture generation, this container only),
  * ``tests/``                 (oracle pinning, when the copy is present),
  * ``bench.py --impl reference`` / ``cpu_baseline`` (CPU baseline on the GPU box).

``oracle/_ref/`` is git-ignored (no reference source enters the history) but is
NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot.

The copy is the reference's own numerics, unchanged.  Only the mechanical
Python-2 -> Python-3 text rewrites stated in SURVEY.md section 8(c) are applied:
  1. ``print X``            -> ``print(X)``         (single-line statements)
  2. ``xrange(``            -> ``range(``
  3. ``.iteritems()``       -> ``.items()``
plus an empty ``matplotlib.pyplot`` stub package (imported at module top of
``code/models/distributions/truncated_normal*.py``, never used on the hot path).
Python-2 implicit relative imports are resolved with ``sys.path`` entries by
``oracle/refimport.py``; nothing in the files is changed for that.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DEFAULT = "/root/reference"
DST = os.path.join(HERE, "_ref", "BNMTF_ARD")

COPY_DIRS = ["code", "tests", os.path.join("data", "toy")]

_PRINT_RE = re.compile(r"^(?P<indent>\s*)(?P<prefix>(?:if [^:]+:\s*)?)print (?P<body>[^(\n].*?)\s*$")
_PRINT_PAREN_TUPLE_RE = re.compile(r"^(?P<indent>\s*)print (?P<body>\(.*\)\s*%\s*.*)$")


def _rewrite_line(line):
    stripped = line.rstrip("\n")
    # `print "..." % (...)`, `print a,b`, `if cond: print "..."`
    m = _PRINT_RE.match(stripped)
    if m and not stripped.lstrip().startswith("#"):
        body = m.group("body")
        if body.endswith("\\"):
            # multi-line print statements only occur inside dead ''' blocks
            return line
        return "%s%sprint(%s)\n" % (m.group("indent"), m.group("prefix"), body)
    m = _PRINT_PAREN_TUPLE_RE.match(stripped)
    if m:
        return "%sprint(%s)\n" % (m.group("indent"), m.group("body"))
    return line


def rewrite_source(text):
    out = []
    for line in text.splitlines(True):
        line = _rewrite_line(line)
        line = line.replace("xrange(", "range(").replace(".iteritems()", ".items()")
        out.append(line)
    return "".join(out)


def build(src=SRC_DEFAULT, dst=DST, quiet=False):
    if not os.path.isdir(src):
        raise FileNotFoundError("reference checkout not found at %s" % src)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    n_files = 0
    for name in os.listdir(src):
        p = os.path.join(src, name)
        if os.path.isfile(p) and name.endswith(".py"):
            with open(p) as f, open(os.path.join(dst, name), "w") as g:
                g.write(rewrite_source(f.read()))
    for d in COPY_DIRS:
        sdir = os.path.join(src, d)
        if not os.path.isdir(sdir):
            continue
        for root, _dirs, files in os.walk(sdir):
            rel = os.path.relpath(root, src)
            os.makedirs(os.path.join(dst, rel), exist_ok=True)
            for fn in files:
                sp = os.path.join(root, fn)
                dp = os.path.join(dst, rel, fn)
                if fn.endswith(".py"):
                    with open(sp, encoding="utf-8") as f:
                        text = f.read()
                    with open(dp, "w", encoding="utf-8") as g:
                        g.write(rewrite_source(text))
                    n_files += 1
                elif fn.endswith((".txt", ".csv")) and os.path.getsize(sp) < 4 << 20:
                    shutil.copyfile(sp, dp)
    # package markers the py2 checkout relied on implicitly
    for d in ["", "data", "tests"]:
        init = os.path.join(dst, d, "__init__.py")
        if os.path.isdir(os.path.dirname(init)) and not os.path.exists(init):
            open(init, "w").close()
    # matplotlib stub (never used on the hot path)
    stub = os.path.join(os.path.dirname(dst), "_stubs", "matplotlib")
    os.makedirs(stub, exist_ok=True)
    open(os.path.join(stub, "__init__.py"), "w").close()
    with open(os.path.join(stub, "pyplot.py"), "w") as f:
        f.write("# empty stub: the reference imports matplotlib.pyplot at module top only\n")
    if not quiet:
        print("reference copy: %d python files -> %s" % (n_files, dst))
    return dst


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
