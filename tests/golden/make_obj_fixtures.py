"""Writes the committed .obj fixtures of the ingest tests (tests/golden/obj/): nothing here comes from the reference.

  quirks.obj  -- hand-written: every line-discipline corner of the reader (comments, CRLF and lone CR, tabs, command
                 followed by several spaces, polygons, negative and zero indices, v with w, exponents, inf/nan,
                 lines Python raises on, corners whose vt is skipped once the triangle lost its vt list, ...)
  torus.obj   -- 32 x 24 torus with texture coordinates and quads, some degenerate faces, a duplicated vertex
                 reference inside one face, an unreferenced vertex; torus.mtl + checker.png (64 x 48) beside it
  fan.obj     -- one vertex shared by 700 triangles with distinct normals plus coplanar duplicates (long incidence list,
                 kept-normal list beyond a warp)
  cube_pm.obj -- axis-aligned cube: face normals full of -0.0 / +0.0 (sign-of-zero behaviour of np.mean)
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "obj")


def quirks():
    L = [
        "# comment line", "", "   ", "#v 9 9 9", " # not a comment: command is empty",
        "v 0 0 1", "v 1 0 1.5", "v  0 1 2 0.5", "v 1e0 1E0 +2.25", "v -.5 .5 3.", "v\t7 7 7", "v 1 2", "v 1 2 x",
        "v 0.1 0.2 1e-3\r", "v 2 2 2\rv 3 3 3", "v 4 4 4 # trailing words break float()", "v inf -Infinity nan",
        "v 1e400 1e-400 4", "v 0x10 1 1", "v 1,5 2 3", "v +1 -2 3 4 5 6",
        "vt 0.25 0.75", "vt 0.5 0.5", "vt 1.5 -0.5", "vt 0 1", "vt nan 0.5", "vt 3e9 -3e9",
        "vn 0 0 1", "vn 0 0 1 0", "vn 1 0", "vn 0 1 0",
        "g group1", "usemtl none", "s off", "mtllib quirks.mtl", "mtllib /nonexistent/abs.mtl",
        "f 1/1/1 2/2/1 3/3/2", "f 1/1/1 2/2/1 3/3/2 4/4/2 5/1/1", "f -1/-1/-1 -2/-2/-2 -3/-3/-1",
        "f 1 2", "f 1/1/1 2/2/2 x/3/3", "f 1/1/1 2/2/2 3/y/3", "f 1.0/1/1 2/2/2 3/3/3", "f 0/0/0 1/1/1 2/2/2",
        "f 1/1/1/9 2/2/2/9 3/3/3/9", "f +1/+1/+1 +2/+2/+2 +3/+3/+3", "f\t1/1/1 2/2/2 3/3/3",
        "f 1//1 2/x/2 3/3/3", "f 4/4/1 5/5/1 6/6/1",
        "f 1/1 2/2 3/3", "f 7/1/zz 8/2/1 9/3/1", "f 7 8 9 10",
        "vt 0.1 0.2", "v 9 8 7",
    ]
    text = "\n".join(L[:14]) + "\r\n" + "\n".join(L[14:]) + "\nv 5 5 5"   # no final newline
    open(os.path.join(OUT, "quirks.obj"), "w", newline="").write(text)
    open(os.path.join(OUT, "quirks.mtl"), "w").write("# material\nnewmtl m\nKd 1 1 1\nmap_Kd missing_first.png\nmap_Kd checker.png\n")


def checker():
    import cv2
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    img[::8] //= 2
    cv2.imwrite(os.path.join(OUT, "checker.png"), img)


def torus(nu=32, nv=24, R=1.0, r=0.35):
    rng = np.random.default_rng(3)
    lines = ["# torus fixture", "mtllib torus.mtl"]
    for i in range(nu):
        for j in range(nv):
            a, b = 2 * np.pi * i / nu, 2 * np.pi * j / nv
            p = ((R + r * np.cos(b)) * np.cos(a), (R + r * np.cos(b)) * np.sin(a), r * np.sin(b) + 3.0)
            lines.append("v %.6f %.6f %.6f" % p)
    lines.append("v 10 10 10")          # never referenced
    for i in range(nu + 1):
        for j in range(nv + 1):
            lines.append("vt %.5f %.5f" % (i / nu * 1.02 - 0.01, j / nv * 1.02 - 0.01))   # a little outside [0,1]
    vid = lambda i, j: (i % nu) * nv + (j % nv) + 1   # noqa: E731
    tid = lambda i, j: i * (nv + 1) + j + 1           # noqa: E731
    for i in range(nu):
        for j in range(nv):
            q = [(vid(i, j), tid(i, j)), (vid(i + 1, j), tid(i + 1, j)), (vid(i + 1, j + 1), tid(i + 1, j + 1)),
                 (vid(i, j + 1), tid(i, j + 1))]
            k = rng.integers(0, 10)
            if k == 0:      # two triangles written separately
                lines.append("f %d/%d %d/%d %d/%d" % (*q[0], *q[1], *q[2]))
                lines.append("f %d/%d %d/%d %d/%d" % (*q[0], *q[2], *q[3]))
            elif k == 1:    # degenerate: a vertex used twice
                lines.append("f %d/%d %d/%d %d/%d %d/%d" % (*q[0], *q[0], *q[2], *q[3]))
            else:
                lines.append("f %d/%d %d/%d %d/%d %d/%d" % (*q[0], *q[1], *q[2], *q[3]))
    open(os.path.join(OUT, "torus.obj"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(OUT, "torus.mtl"), "w").write("newmtl t\nmap_Kd checker.png\n")


def fan(n=700):
    rng = np.random.default_rng(11)
    lines = ["v 0 0 2"]
    for i in range(n + 1):
        a = 2 * np.pi * i / n
        lines.append("v %.7f %.7f %.7f" % (np.cos(a), np.sin(a), 2.0 + 0.3 * np.sin(5 * a) + 0.01 * rng.standard_normal()))
    for i in range(n):
        lines.append("f 1 %d %d" % (i + 2, i + 3))
    # coplanar duplicates of a few fan triangles (their normals repeat exactly) and a flat ring around vertex 1
    for i in (0, 5, 333, 699):
        lines.append("f 1 %d %d" % (i + 2, i + 3))
    lines += ["v 1 0 2", "v 0 1 2", "v -1 0 2", "v 0 -1 2"]
    b = n + 3
    for k in range(4):
        lines.append("f 1 %d %d" % (b + k, b + (k + 1) % 4))
    open(os.path.join(OUT, "fan.obj"), "w").write("\n".join(lines) + "\n")


def cube():
    v = [(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (2, 4)]
    lines = ["v %d %d %d" % p for p in v]
    quads = [(1, 2, 4, 3), (5, 7, 8, 6), (1, 5, 6, 2), (3, 4, 8, 7), (1, 3, 7, 5), (2, 6, 8, 4)]
    for q in quads:
        lines.append("f %d %d %d %d" % q)
    lines.append("f 1 1 2")   # zero normal: joins both vertices' lists every time it is met
    open(os.path.join(OUT, "cube_pm.obj"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    quirks()
    checker()
    torus()
    fan()
    cube()
    print(sorted(os.listdir(OUT)))
