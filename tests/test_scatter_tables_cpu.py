"""The compile-time scatter tables of the fused assembly kernel (tc-viml_b200/csrc/assemble2.cuh, `make_tables`) checked on
the CPU: the constexpr code between the "scatter tables" banner and the `g_tables` definition is plain C++17, so it is cut out
of the header where it lies, compiled with g++ and its tables dumped.  The test then replays what the kernel does with them —
Gram matrix U^T U of the 16 factor columns per (i, j) segment, segment / anchor / extrinsic / line flushes into the block-upper
accumulator, expansion with the mirror rules — in numpy and compares the result with the dense J^T J / J^T r the reference
builds (marginalization_factor.cpp:141-172) from the same Jacobians.  No GPU, no product code executed."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CUH = os.path.join(ROOT, "tc-viml_b200", "csrc", "assemble2.cuh")
KINDS = ["LL", "HH", "LH", "LE", "HE", "EE", "BL", "BH", "BE"]


@pytest.fixture(scope="module")
def tables(tmp_path_factory):
    src = open(CUH).read()
    a = src.index("enum Kind {")
    b = src.index("__device__ const Tables g_tables")
    body = src[a:b]
    tmp = tmp_path_factory.mktemp("tables")
    cpp = tmp / "dump.cpp"
    cpp.write_text(
        "#include <cstdint>\n#include <cstdio>\n" + body +
        "static void dump(const char* n, const uint32_t* t, int len, bool last) {\n"
        "  std::printf(\"\\\"%s\\\": [\", n);\n"
        "  for (int i = 0; i < len; ++i) std::printf(\"%s%u\", i ? \",\" : \"\", t[i]);\n"
        "  std::printf(\"]%s\\n\", last ? \"\" : \",\");\n}\n"
        "int main() {\n  constexpr Tables T = make_tables();\n  std::printf(\"{\\n\");\n"
        "  dump(\"seg\", T.seg, sizeof(T.seg) / 4, false); dump(\"lo\", T.lo, sizeof(T.lo) / 4, false);\n"
        "  dump(\"ee\", T.ee, sizeof(T.ee) / 4, false); dump(\"line\", T.line, sizeof(T.line) / 4, true);\n"
        "  std::printf(\"}\\n\");\n}\n")
    exe = tmp / "dump"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", str(exe), str(cpp)])
    return json.loads(subprocess.check_output([str(exe)]))


def decode(words):
    """tab_entry: valid<<31 | neg<<18 | off<<12 | kind<<8 | src (tile*64 + row*8 + col)"""
    out = []
    for w in words:
        if not (w >> 31):
            continue
        out.append((w & 0xff, KINDS[(w >> 8) & 15], (w >> 12) & 63, -1.0 if (w >> 18) & 1 else 1.0))
    return out


def patch_of(G):
    """The 192-double patch: the three 8x8 tiles (rows 0-7 x cols 0-7, rows 0-7 x cols 8-15, rows 8-15 x cols 8-15)."""
    return np.concatenate([G[0:8, 0:8].ravel(), G[0:8, 8:16].ravel(), G[8:16, 8:16].ravel()])


def blk(br, bc, NB):
    return br * NB - (br * (br - 1)) // 2 + (bc - br)


def test_table_shapes(tables):
    seg, lo, ee, line = (decode(tables[k]) for k in ("seg", "lo", "ee", "line"))
    assert len(seg) == 96 and len(lo) == 63 and len(ee) == 27 and len(line) == 27   # 3 / 2 / 1 / 1 rounds of 32 lanes
    assert [k for _, k, _, _ in seg].count("LH") == 33                               # symmetric 3x3 as upper triangle
    for t in (seg, lo, ee, line):
        dests = [(k, o) for _, k, o, _ in t]
        assert len(set(dests)) == len(dests), "a destination is fed by exactly one Gram entry"
    # the table words sit densely from index 0 (a round = 32 consecutive words)
    for name, n in (("seg", 96), ("lo", 63), ("ee", 27), ("line", 27)):
        w = tables[name]
        assert all(x >> 31 for x in w[:n]) and not any(x >> 31 for x in w[n:])


def test_gram_scatter_equals_dense_normal_equations(tables):
    rng = np.random.default_rng(5)
    seg, lo_t, ee_t, line_t = (decode(tables[k]) for k in ("seg", "lo", "ee", "line"))
    P = 5
    NB = P + 1
    D = 6 * NB
    nblk = NB * (NB + 1) // 2
    boff = nblk * 36
    Hc = np.zeros(boff + D)
    Href = np.zeros((D, D))
    bref = np.zeros(D)

    def base(kind, lo, hi):
        e = NB - 1
        rows = dict(LL=(lo, lo), HH=(hi, hi), LH=(lo, hi), LE=(lo, e), HE=(hi, e), EE=(e, e))
        if kind in rows:
            return blk(*rows[kind], NB) * 36
        return boff + 6 * dict(BL=lo, BH=hi, BE=e)[kind]

    def flush(table, G, lo, hi):
        p = patch_of(G)
        for src, kind, off, sign in table:
            Hc[base(kind, lo, hi) + off] += sign * p[src]

    E = np.zeros((16, 16))
    # point factors: anchors i < j as in the reference (est.cpp:1735-1770), several factors per (i, j) segment
    for i in range(P - 1):
        R = np.zeros((16, 16))
        for j in range(i + 1, P):
            G = np.zeros((16, 16))
            for _ in range(int(rng.integers(1, 4))):
                A = rng.standard_normal((2, 3))            # translation block of pose i; pose j's is -A
                Xr, Br = rng.standard_normal((2, 3)), rng.standard_normal((2, 3))
                Z, r = rng.standard_normal((2, 6)), rng.standard_normal((2, 1))
                U = np.hstack([A, Xr, Br, Z, r])           # the factor's 16 distinct columns (lo role = i)
                G += U.T @ U
                J = np.zeros((2, D))
                J[:, 6 * i:6 * i + 6] = np.hstack([A, Xr])
                J[:, 6 * j:6 * j + 6] = np.hstack([-A, Br])
                J[:, 6 * P:6 * P + 6] = Z
                Href += J.T @ J
                bref += (J.T @ r).ravel()
            flush(seg, G, i, j)                             # end of the segment
            R += G
            E += G
        Rlo = R.copy()
        Rlo[8:, 8:] = 0.0                                   # tile 2 holds no lo-role entry
        flush(lo_t, Rlo, i, i)                              # anchor change
    Ee = np.zeros((16, 16))
    Ee[8:, 8:] = E[8:, 8:]
    flush(ee_t, Ee, 0, 0)                                   # once per warp
    # line factors of every frame: columns 0..5 = J, 6 = residual
    for p in range(P):
        G = np.zeros((16, 16))
        for _ in range(3):
            Jl, r = rng.standard_normal((2, 6)), rng.standard_normal((2, 1))
            U = np.zeros((2, 16))
            U[:, :6], U[:, 6:7] = Jl, r
            G += U.T @ U
            Href[6 * p:6 * p + 6, 6 * p:6 * p + 6] += Jl.T @ Jl
            bref[6 * p:6 * p + 6] += (Jl.T @ r).ravel()
        flush(line_t, G, p, p)

    # expansion: upper blocks as they are, mirrored blocks transposed, diagonal blocks symmetrised, and the symmetric
    # top-left 3x3 of a pose-pose block completed from its upper triangle
    H = np.zeros((D, D))
    for br in range(NB):
        for bc in range(br, NB):
            B = Hc[blk(br, bc, NB) * 36:][:36].reshape(6, 6).copy()
            if br == bc:
                B = np.triu(B) + np.triu(B, 1).T
            elif bc < NB - 1:
                for a in range(3):
                    for b in range(a):
                        B[a, b] = B[b, a]
            H[6 * br:6 * br + 6, 6 * bc:6 * bc + 6] = B
            H[6 * bc:6 * bc + 6, 6 * br:6 * br + 6] = B.T
    bp = Hc[boff:boff + D]
    assert np.allclose(H, Href, rtol=0, atol=1e-12 * np.abs(Href).max())
    assert np.allclose(bp, bref, rtol=0, atol=1e-12 * np.abs(bref).max())
