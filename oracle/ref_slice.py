"""Build-time slicer for oracle/_ref (TEST INFRASTRUCTURE ONLY).

estimator.cpp cannot be compiled as a translation unit here (it needs Ceres problems, OpenCV and the initialisation code), but
the four association functions of SURVEY 8a (UpdateLinesInFoV, CalAngleDist, CalEulerDist, LineCorrespondenceInFrame) only touch
a handful of Estimator members.  This script reads the reference file WHERE IT LIES, finds those member-function definitions by
their signature, and writes their text, unmodified, to a scratch include that oracle/ref_estimator.cpp compiles against a minimal
`Estimator` declaration.  The scratch file is a build intermediate under oracle/_ref/ (git-ignored) and the Makefile deletes it
after the compile: no reference source is committed or shipped.

usage: python ref_slice.py <path to estimator.cpp> <output include> <function name> [...]"""
import re
import sys


def slice_function(text, name):
    """Text of the definition `<ret> Estimator::<name>(...) { ... }`, from the start of its first line to the closing brace."""
    m = re.search(r"^[^\n;{}#]*\bEstimator::" + re.escape(name) + r"\s*\(", text, re.M)
    if not m:
        raise SystemExit(f"ref_slice: Estimator::{name} not found")
    i = text.index("{", m.end())
    if ";" in text[m.end():i]:
        raise SystemExit(f"ref_slice: Estimator::{name}: declaration, not a definition")
    depth, k = 0, i
    in_line = in_block = in_str = in_chr = False
    while k < len(text):
        c, n = text[k], text[k + 1:k + 2]
        if in_line:
            in_line = c != "\n"
        elif in_block:
            if c == "*" and n == "/":
                in_block, k = False, k + 1
        elif in_str:
            if c == "\\":
                k += 1
            elif c == '"':
                in_str = False
        elif in_chr:
            if c == "\\":
                k += 1
            elif c == "'":
                in_chr = False
        elif c == "/" and n == "/":
            in_line = True
        elif c == "/" and n == "*":
            in_block = True
        elif c == '"':
            in_str = True
        elif c == "'":
            in_chr = True
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[m.start():k + 1], text.count("\n", 0, m.start()) + 1
        k += 1
    raise SystemExit(f"ref_slice: Estimator::{name}: unbalanced braces")


def main():
    src, out, names = sys.argv[1], sys.argv[2], sys.argv[3:]
    text = open(src, encoding="utf-8", errors="replace").read()
    with open(out, "w") as f:
        for name in names:
            body, line = slice_function(text, name)
            f.write(f'#line {line} "{src}"\n{body}\n\n')


if __name__ == "__main__":
    main()
