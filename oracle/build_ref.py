#!/usr/bin/env python3
"""Build the UNMODIFIED-ALGORITHM reference (casutton/bayes-qnet, Python 2 + Cython 0.2x) as an
importable Python 3.12 / Cython 3 package under ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bayes_qnet_b200/`` may import or execute this.

The reference cannot be imported as shipped (Python 2 syntax, Py2-only stdlib semantics, two modules
missing from the checkout).  This script reads the sources where they lie (``/root/reference/src``),
applies a purely *mechanical* Py2->Py3 patch (no algorithmic change: print statements, ``except X, e``,
``iteritems``, ``cmp=``, dict views, ``yaml.load`` Loader, heap ordering of mixed tuples) to a scratch copy
in ``oracle/_ref/`` (git-ignored; never committed) and compiles the eight Cython extensions exactly as
``src/setup.py:15-27`` lists them.  The result is what ``oracle/make_golden.py`` imports to generate
``tests/golden/*.npz`` and what ``bench.py --impl reference`` times on the host cores.

Usage:  python oracle/build_ref.py [--src /root/reference/src] [--force]
"""
import argparse
import glob
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PYX = ["cninq", "misc", "pwfun", "sampling", "arrivals", "queues", "qnet", "distributions"]
PY = ["qnetu", "estimation", "hmm", "netutils", "ninqueue", "qstats", "utils", "ivltreap", "stupidlp"]
EXTRA = ["randomkit.c", "randomkit.h", "cdist.c", "cdist.h", "arrv_true.txt"]

CMP_DEF = "def cmp(a, b): return (a > b) - (a < b)\n"
CMP_KEY = "key=__import__('functools').cmp_to_key(%s)"

HEAP_SHIM = '''from heapq import heappush as _hpush, heappop as _hpop
def _p2k(x):
    # Python 2 ordered None < numbers < everything else; the simulator's heap tuples mix all three
    if x is None: return (0, 0)
    if isinstance(x, (int, float)): return (1, x)
    return (2, x)
def heappush(h, t): _hpush(h, (tuple([_p2k(x) for x in t]), t))
def heappop(h): return _hpop(h)[1]
'''

MYTIME = '''import time
class timeit:
    def __init__(self): self.t0 = self.last = time.time()
    def tick(self, msg=""):
        now = time.time(); print("%s: %.4f s" % (msg, now - self.last)); self.last = now
    def total(self, msg=""):
        now = time.time(); print("%s: %.4f s" % (msg, now - self.t0)); return now - self.t0
'''

SETUP = '''from setuptools import setup, Extension
from Cython.Build import cythonize
exts = [Extension("cninq", ["cninq.pyx"]), Extension("misc", ["misc.pyx"]), Extension("pwfun", ["pwfun.pyx"]),
        Extension("sampling", ["sampling.pyx", "randomkit.c", "cdist.c"]), Extension("arrivals", ["arrivals.pyx"]),
        Extension("queues", ["queues.pyx"]), Extension("qnet", ["qnet.pyx"]),
        Extension("distributions", ["distributions.pyx"])]
setup(name="bayes_qnet_ref", ext_modules=cythonize(exts, language_level=2, quiet=True))
'''


def _split_comment(line):
    """Split a trailing '# comment' (outside string literals) off a source line."""
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q and line[i - 1] != '\\':
                q = None
        elif ch in "\"'":
            q = ch
        elif ch == '#':
            return line[:i].rstrip() + "\n", "  " + line[i:].rstrip("\n")
    return line, ""


def _fix_print(line):
    line, cmt = _split_comment(line)
    out = _fix_print0(line)
    return out.rstrip("\n") + cmt + "\n" if cmt else out


def _fix_print0(line):
    m = re.match(r'^(\s*)print\s*$', line)
    if m:
        return m.group(1) + "print()\n"
    m = re.match(r'^(\s*)print\s+(?!\()(.*?)\s*$', line)
    if not m:
        m2 = re.match(r'^(\s*)print\s*(\(.*\))\s*,\s*$', line)
        if m2:
            return "%sprint(%s, end=' ')\n" % (m2.group(1), m2.group(2))
        return line
    ind, body = m.group(1), m.group(2).rstrip()
    if body.endswith(','):
        return "%sprint(%s end=' ')\n" % (ind, body)
    return "%sprint(%s)\n" % (ind, body)


def _fix_inline_print(line):
    # "if x: print a, b" forms
    if line.lstrip().startswith("#"):
        return line
    line0, cmt = _split_comment(line)
    if cmt and re.match(r'^(.*?:\s*)print\s+(?!\()', line0):
        return _fix_inline_print(line0).rstrip("\n") + cmt + "\n"
    m = re.match(r'^(.*?:\s*)print\s+(?!\()(.*?)\s*$', line)
    if m and not line.lstrip().startswith("#") and "print(" not in m.group(1):
        body = m.group(2)
        if body.endswith(','):
            return "%sprint(%s end=' ')\n" % (m.group(1), body)
        return "%sprint(%s)\n" % (m.group(1), body)
    return line


def patch_py(text):
    """Mechanical Py2->Py3 for the plain-python modules."""
    out = []
    for line in text.replace("\t", "        ").splitlines(True):
        s = line.lstrip()
        if re.match(r'print(\s|$)', s):
            line = _fix_print(line)
        else:
            line = _fix_inline_print(line)
        line = re.sub(r'except\s+([\w\.]+)\s*,\s*(\w+)\s*:', r'except \1 as \2:', line)
        line = line.replace(".iteritems()", ".items()").replace(".iterkeys()", ".keys()").replace("xrange", "range")
        line = re.sub(r'raise\s+"(.*)"\s*%\s*(.*)$', r'raise Exception("\1" % \2)', line)
        line = re.sub(r'yaml\.load\s*\((\w+)\)', r'yaml.load(\1, Loader=yaml.Loader)', line)
        out.append(line)
    return "".join(out)


def _sub(text, old, new, count=None, must=True):
    n = text.count(old)
    if must and n == 0:
        raise RuntimeError("patch anchor not found: %r" % old)
    return text.replace(old, new) if count is None else text.replace(old, new, count)


def patch_module(name, text):
    if name == "arrivals.pyx":
        text = CMP_DEF + text
        text = _sub(text, "import mytime\n", "")
        text = _sub(text, "    def __cmp__ (self, other):",
                    "    def __lt__ (self, other): return self.__cmp__(other) < 0\n"
                    "    def __le__ (self, other): return self.__cmp__(other) <= 0\n"
                    "    def __gt__ (self, other): return self.__cmp__(other) > 0\n"
                    "    def __ge__ (self, other): return self.__cmp__(other) >= 0\n"
                    "    def __cmp__ (self, other):")
        text = _sub(text, "evtl.sort(cmp=cmp_by_arrival)", "evtl.sort(%s)" % (CMP_KEY % "cmp_by_arrival"))
        text = _sub(text, "evtl.sort(cmp=cmp_by_departure)", "evtl.sort(%s)" % (CMP_KEY % "cmp_by_departure"))
        text = _sub(text, "return self.events.values().__iter__()", "return iter(list(self.events.values()))")
        text = _sub(text, "        return self.events.values()\n", "        return list(self.events.values())\n")
        text = _sub(text, "def all_tasks (self): return self.byt.keys()", "def all_tasks (self): return list(self.byt.keys())")
        text = _sub(text, "        for tid in self.byt.keys():", "        for tid in list(self.byt.keys()):")
        text = text.replace(".iteritems()", ".items()").replace(".iterkeys()", ".keys()")
        # capacity knob, default unchanged: the reference sizes its per-queue sorted arrays for at most BIG_N = 10000
        # events per queue (arrivals.pyx:237, cninq.pyx:146-147) and raises beyond; BASELINE configs 3 and 4 have
        # 20 000 / 50 000 events per queue, so the config-size golden vectors and CPU baselines set BQ_REF_BIG_N
        text = _sub(text, "BIG_N = 10000\n", "import os as _os\nBIG_N = int(_os.environ.get('BQ_REF_BIG_N', '10000'))\n")
    elif name == "cninq.pyx":
        text = CMP_DEF + text
        text = _sub(text, "lst.sort(cmp_knots)", "lst.sort(%s)" % (CMP_KEY % "cmp_knots"))
    elif name == "qnet.pyx":
        text = _sub(text, "from heapq import heappush, heappop\n", HEAP_SHIM)
        text = text.replace(".iteritems()", ".items()")
        text = re.sub(r'except\s+([\w\.]+)\s*,\s*(\w+)\s*:', r'except \1 as \2:', text)
    elif name == "queues.pyx":
        text = _sub(text, "eid = changed.keys()[0]", "eid = list(changed.keys())[0]")
    elif name == "pwfun.pyx":
        text = re.sub(r'return map ?\((.*)\)\n', r'return list(map(\1))\n', text)
    elif name == "distributions.pyx":
        text = _sub(text, "numpy.mean (map (numpy.log, data))", "numpy.mean (list(map (numpy.log, data)))")
        text = text.replace("\t", "        ")
    elif name == "sampling.pyx":
        # DEBUG=1 in the checkout prints every slice step (sampling.pyx:17,351)
        text = re.sub(r'^DEBUG\s*=\s*1', 'DEBUG = 0', text, flags=re.M)
    elif name == "qnetu.py":
        text = re.sub(r'^(\s*)map \(add_state, (.*)\)$', r'\1list(map(add_state, \2))', text, flags=re.M)
        text = re.sub(r'^(\s*)map \(add_queue, (.*)\)$', r'\1list(map(add_queue, \2))', text, flags=re.M)
        text = re.sub(r'= map \((sname2id|qname2id)\.get, (.*)\)$', r'= list(map(\1.get, \2))', text, flags=re.M)
        text = text.replace("yaml.load (yaml_text)", "yaml.load(yaml_text, Loader=yaml.Loader)")
        text = text.replace("yaml.load (text)", "yaml.load(text, Loader=yaml.Loader)")
        text = text.replace("yaml.load (stream)", "yaml.load(stream, Loader=yaml.Loader)")
    elif name == "netutils.py":
        text = text.replace("numpy.sum(map(numpy.log, data))", "numpy.sum(list(map(numpy.log, data)))")
    elif name == "ninqueue.py":
        text = CMP_DEF + text
        text = _sub(text, "lst.sort(cmp_knots)", "lst.sort(%s)" % (CMP_KEY % "cmp_knots"))
        text = _sub(text, "times = knots.keys()", "times = list(knots.keys())")
        text = _sub(text, "lst = knots.items()", "lst = list(knots.items())")
    return text


def build(src, force=False, py_only=False):
    marker = os.path.join(OUT, ".built")
    if py_only:
        for m in PY:
            text = patch_module(m + ".py", patch_py(open(os.path.join(src, m + ".py")).read()))
            open(os.path.join(OUT, m + ".py"), "w").write(text)
        return OUT
    if os.path.exists(marker) and not force:
        print("oracle/_ref already built (use --force to rebuild)")
        return OUT
    if not os.path.isdir(src):
        raise SystemExit("reference sources not found at %s" % src)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for m in PYX:
        for ext in (".pyx", ".pxd"):
            p = os.path.join(src, m + ext)
            if os.path.exists(p):
                text = open(p).read()
                text = patch_module(m + ext, text)
                open(os.path.join(OUT, m + ext), "w").write(text)
    for m in PY:
        p = os.path.join(src, m + ".py")
        text = patch_py(open(p).read())
        text = patch_module(m + ".py", text)
        open(os.path.join(OUT, m + ".py"), "w").write(text)
    for f in EXTRA:
        shutil.copy(os.path.join(src, f), os.path.join(OUT, f))
    exdir = os.path.join(os.path.dirname(src), "examples", "models")
    if os.path.isdir(exdir):
        os.makedirs(os.path.join(OUT, "examples"), exist_ok=True)
        for f in glob.glob(os.path.join(exdir, "*.yml")):
            shutil.copy(f, os.path.join(OUT, "examples", os.path.basename(f)))
    open(os.path.join(OUT, "mytime.py"), "w").write(MYTIME)
    open(os.path.join(OUT, "setup3.py"), "w").write(SETUP)
    log = open(os.path.join(OUT, "build.log"), "w")
    r = subprocess.run([sys.executable, "setup3.py", "build_ext", "--inplace"], cwd=OUT, stdout=log,
                       stderr=subprocess.STDOUT)
    log.close()
    if r.returncode != 0:
        sys.stderr.write(open(os.path.join(OUT, "build.log")).read()[-4000:])
        raise SystemExit("reference build failed; see oracle/_ref/build.log")
    # the generated .c files and build tree are large and not needed at run time
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    for m in PYX:
        c = os.path.join(OUT, m + ".c")
        if os.path.exists(c):
            os.remove(c)
    open(marker, "w").write("ok\n")
    print("built reference into", OUT)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference/src")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--py-only", action="store_true", help="re-patch only the plain .py modules")
    a = ap.parse_args()
    build(a.src, a.force, a.py_only)
