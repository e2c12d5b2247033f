"""A second, independent restatement of the reference's intersection code in plain Python floats.

TEST INFRASTRUCTURE.  Written from the Rust sources (file:line below, relative to the reference checkout),
not from oracle/raymond_oracle.cpp, so that a transcription slip in one of the two shows up as a bit
difference between them (tests/test_oracle.py).  Python floats are IEEE f64 and CPython never fuses a*b+c,
which is exactly the arithmetic rustc emits.  Pure-Python loops: small cases only.
"""
from __future__ import annotations

import math
import struct

F_MAX = 1.7976931348623157e308          # core/src/math.rs:20


# ---- cgmath 0.17 vector semantics
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def neg(a): return (-a[0], -a[1], -a[2])
def scale(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def dot(a, b): return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def fdiv(a, b):
    """IEEE division (Python raises on /0)."""
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        neg_ = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg_ else math.inf
    return a / b


def fmin(a, b):
    """f64::min: the non-NaN operand."""
    if a != a: return b
    if b != b: return a
    return a if a < b else b


def fmax(a, b):
    if a != a: return b
    if b != b: return a
    return a if a > b else b


def cast_i32(v):
    """cgmath cast::<i32>() = NumCast: Some(trunc toward zero) when in range, else None."""
    if v != v or not (-2147483649.0 < v < 2147483648.0):
        return None
    return int(v)


def cast_usize(v):
    if v != v or not (-1.0 < v < 18446744073709551616.0):
        return None
    return int(v)


def as_usize(v):
    """Rust `f64 as usize`: saturating, NaN -> 0."""
    if v != v or v <= 0.0:
        return 0
    if v >= 18446744073709551616.0:
        return 2**64 - 1
    return int(v)


# ---- primitives
def sphere_intersects(origin, radius, o, d):            # primitives/sphere.rs:11-27
    c = sub(origin, o)
    t = dot(c, d)
    q = sub(c, scale(d, t))          # t * ray.direction: commutative per component
    p = dot(q, q)
    if p > radius * radius:
        return None
    t -= math.sqrt(radius * radius - p)
    if t <= 0.0:
        return None
    return t


def plane_intersects(origin, normal, o, d):             # primitives/plane.rs:11-24
    denom = dot(normal, neg(d))
    if denom > 1e-6:
        p0l0 = sub(origin, o)
        t = dot(p0l0, neg(normal)) / denom
        if t >= 0.0:
            return t
    return None


def aabb_intersects(bmin, bmax, o, d):                   # primitives/aabb.rs:10-31
    inv = (fdiv(1.0, d[0]), fdiv(1.0, d[1]), fdiv(1.0, d[2]))
    t1 = (bmin[0] - o[0]) * inv[0]
    t2 = (bmax[0] - o[0]) * inv[0]
    tmin, tmax = fmin(t1, t2), fmax(t1, t2)
    for i in (1, 2):
        t1 = (bmin[i] - o[i]) * inv[i]
        t2 = (bmax[i] - o[i]) * inv[i]
        tmin = fmax(tmin, fmin(t1, t2))
        tmax = fmin(tmax, fmax(t1, t2))
    if not (tmax > fmax(tmin, 0.0)):
        return None
    return tmin


def triangle_intersects(v0, v1, v2, o, d):               # primitives/triangle.rs:11-44
    EPS = 0.00000001
    e1 = sub(v1, v0)
    e2 = sub(v2, v0)
    h = cross(d, e2)
    a = dot(e1, h)
    if a < EPS and a > -EPS:
        return None
    f = 1.0 / a
    s = sub(o, v0)
    u = f * dot(s, h)
    if u < 0.0 or u > 1.0:
        return None
    q = cross(s, e1)
    v = f * dot(d, q)
    if v < 0.0 or u + v > 1.0:
        return None
    t = f * dot(e2, q)
    if t > EPS:
        return t
    return None


def find_bounds(points):                                 # triangle.rs:70-84, mesh.rs:123-140
    mn = [125125.0, 1251251.0, 12512512.0]
    mx = [-123125.0, -125123.0, -512123.0]
    for p in points:
        for i in range(3):
            mn[i] = fmin(mn[i], p[i])
            mx[i] = fmax(mx[i], p[i])
    return tuple(mn), tuple(mx)


class GridOOB(Exception):
    """naive_cells[..] index out of bounds (acc_grid.rs:61 panics)."""


class Grid:
    """AccGrid::build_from_mesh (acc_grid.rs:36-83) over a list of triangles ((v0, v1, v2) position triples)."""

    def __init__(self, tris):
        self.tris = tris
        self.bmin, self.bmax = find_bounds([p for t in tris for p in t])
        size = sub(self.bmax, self.bmin)
        volume = abs(size[0] * size[1] * size[2])                          # acc_grid.rs:6-17
        density = math.pow((3.0 * float(len(tris))) / volume, 1.0 / 3.0)
        self.res = tuple(as_usize(abs(size[i]) * density) for i in range(3))
        self.cell = tuple(size[i] / float(self.res[i]) for i in range(3))
        n_cells = self.res[0] * self.res[1] * self.res[2]
        naive = [[] for _ in range(n_cells)]
        for index, t in enumerate(tris):
            bmn, bmx = find_bounds(t)
            cmin = [cast_usize(fdiv(bmn[i] - self.bmin[i], self.cell[i])) for i in range(3)]
            cmax = [cast_usize(fdiv(bmx[i] - self.bmin[i], self.cell[i])) for i in range(3)]
            assert None not in cmin and None not in cmax, "Failed to cast cell bounds to usize"
            for i in range(3):
                cmin[i] = min(max(cmin[i], 0), self.res[i] - 1)
                cmax[i] = min(max(cmax[i], 0), self.res[i] - 1)
            for z in range(cmin[2], cmax[2] + 1):
                for y in range(cmin[1], cmax[1] + 1):
                    for x in range(cmin[0], cmax[0] + 1):
                        idx = x + self.res[0] * (y + z * self.res[2])    # sic: res.z
                        if idx >= n_cells:
                            raise GridOOB(index)
                        naive[idx].append(index)
        self.cells = []
        self.table = []
        for c in naive:
            self.cells.append(len(self.table))
            self.table.append(len(c))
            self.table.extend(c)

    def intersects(self, o, d):                            # acc_grid.rs:89-185
        tmin = aabb_intersects(self.bmin, self.bmax, o, d)
        if tmin is None:
            return None
        outer = add(o, scale(d, tmin))
        start = sub(o, self.bmin)
        cell = [cast_i32(fdiv(start[i], self.cell[i])) for i in range(3)]
        if None in cell:
            return None                                    # the reference panics (unwrap)
        if cell[0] < 0 or cell[1] < 0 or cell[2] < 0:
            start = sub(outer, self.bmin)
            cell = [cast_i32(fdiv(start[i], self.cell[i])) for i in range(3)]
            if None in cell:
                return None
        if any(x != x for x in d):
            return None                                    # signum(NaN).cast().unwrap() panics
        step = [(-1 if math.copysign(1.0, d[i]) < 0 else 1) for i in range(3)]
        t_delta = [fdiv(-self.cell[i] if d[i] < 0.0 else self.cell[i], d[i]) for i in range(3)]
        t_max = [fdiv((float(cell[i] + (0 if d[i] < 0.0 else 1)) * self.cell[i]) - start[i], d[i]) for i in range(3)]
        ncells = len(self.cells)
        while True:
            x, y, z = (c & (2**64 - 1) for c in cell)       # i32 as usize sign-extends
            idx = (x + self.res[0] * (y + z * self.res[2])) & (2**64 - 1)
            if idx >= ncells:
                return None
            base = self.cells[idx]
            count = self.table[base]
            closest = 5712515.0
            hit = None
            for i in range(1, count + 1):
                ti = self.table[base + i]
                t = triangle_intersects(*self.tris[ti], o, d)
                if t is not None and t < closest:
                    closest = t
                    hit = (ti, t)
            if hit is not None:
                return hit
            if t_max[0] < t_max[1]:
                a = 0 if t_max[0] < t_max[2] else 2
            else:
                a = 1 if t_max[1] < t_max[2] else 2
            cell[a] += step[a]
            if cell[a] >= self.res[a] or cell[a] < 0:
                return None
            t_max[a] += t_delta[a]


def scene_intersect(objects, o, d):                        # core/src/scene.rs:54-74
    """objects: list of ("sphere", origin, radius) | ("plane", origin, normal) | ("grid", Grid).
    Returns (object index, subobject index, distance) or None."""
    closest = F_MAX
    best = None
    for i, ob in enumerate(objects):
        if ob[0] == "sphere":
            t = sphere_intersects(ob[1], ob[2], o, d)
            h = None if t is None else (0, t)
        elif ob[0] == "plane":
            t = plane_intersects(ob[1], ob[2], o, d)
            h = None if t is None else (0, t)
        else:
            h = ob[1].intersects(o, d)
        if h is not None and h[1] < closest:
            closest = h[1]
            best = (i, h[0], h[1])
    return best


def primary_ray(x, y, cam, jx=0.0, jy=0.0):               # src/trace.rs:322-333 (jitter terms given)
    width, height = float(cam["width"]), float(cam["height"])
    aspect = width / height
    xf = float(x) + jx
    yf = float(y) + jy
    th = math.tan(cam["fov_vert"] / 2.0 * 3.14159265358979323846 / 180.0)
    px = (2.0 * ((xf + 0.5) / width) - 1.0) * th * aspect
    py = (1.0 - 2.0 * ((yf + 0.5) / height)) * th
    v = (px, py, 1.0)
    inv = 1.0 / math.sqrt(dot(v, v))
    return tuple(cam.get("position", (0.0, 0.0, 0.0))), scale(v, inv)


def tile_layout(W, H, tw, th):                             # src/trace.rs:142-173
    out = []
    x = y = 0
    while True:
        mx, my = min(x + tw, W), min(y + th, H)
        out.append((x, y, mx - x, my - y))
        y += th
        if y >= H:
            y = 0
            x += tw
        if x >= W:
            break
    return out


def bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]
