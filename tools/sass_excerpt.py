#!/usr/bin/env python
"""SASS evidence for profiles/: opcode histogram + an excerpt of every function of liblsp_b200.so (or another
object) whose name matches a regex.   python tools/sass_excerpt.py 'k_int_peak' [lib] [excerpt_lines]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"/\*[0-9a-f]+\*/", "", line).strip().rstrip(";").strip()
            if ins:
                body.append(ins)
    if name:
        yield name, body


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def opcode(ins):
    t = ins.split()
    if t[0].startswith("@"):
        t = t[1:]
    op = t[0]
    # carry-out predicate makes a different issue form: IMAD.WIDE.U32 Rd, P0, ... / IADD3 Rd, P0, P1, ...
    if op.startswith("IMAD.WIDE") and len(t) > 2 and re.match(r"P\d", t[2]):
        op += " (carry-out)"
    return op


def main():
    pat = re.compile(sys.argv[1])
    lib = Path(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] else ROOT / "linea-stark-prover_b200" / "liblsp_b200.so"
    n_ex = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    for name, body in functions(lib):
        d = demangle(name)
        if not pat.search(d):
            continue
        hist = collections.Counter(opcode(i) for i in body)
        print(f"== {d}   ({len(body)} instructions)")
        print("   " + ", ".join(f"{k} x{v}" for k, v in hist.most_common(14)))
        # excerpt: the densest run of multiplies
        idx = [i for i, x in enumerate(body) if "IMAD.WIDE" in x or "IMAD.HI" in x]
        start = idx[len(idx) // 2] if idx else 0
        for x in body[start:start + n_ex]:
            print("      " + x)
        print()


if __name__ == "__main__":
    main()
