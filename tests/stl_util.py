"""Small closed triangle meshes written as binary STL files for the tests (no reference fixture is copied)."""
import struct

import numpy as np


def write_binary_stl(path, triangles):
    tri = np.asarray(triangles, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(b"s3b200 test mesh".ljust(80, b" "))
        f.write(struct.pack("<I", tri.shape[0]))
        for t in tri:
            n = np.cross(t[1] - t[0], t[2] - t[0])
            nn = np.linalg.norm(n)
            n = n / nn if nn > 0 else n
            f.write(struct.pack("<3f", *n))
            for v in t:
                f.write(struct.pack("<3f", *v))
            f.write(struct.pack("<H", 0))


def cube_triangles(lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0],
                  [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]])
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3)]
    tris = []
    for a, b, c, d in quads:
        tris.append([v[a], v[b], v[c]])
        tris.append([v[a], v[c], v[d]])
    return np.array(tris)


def icosphere_triangles(subdivisions=2, radius=1.0, center=(0.0, 0.0, 0.0)):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(v, dtype=np.float64) / np.linalg.norm(v) for v in verts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
             (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
             (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, new_faces = {}, []

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = verts[i] + verts[j]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    v = np.array(verts) * radius + np.array(center)
    return np.array([[v[a], v[b], v[c]] for a, b, c in faces])


def torus_triangles(n_major=48, n_minor=24, R=1.0, r=0.35, center=(0.0, 0.0, 0.0)):
    """Closed torus (axis z), parametric grid split into triangles; vertices lie ON the analytic torus."""
    u = np.linspace(0.0, 2.0 * np.pi, n_major, endpoint=False)
    v = np.linspace(0.0, 2.0 * np.pi, n_minor, endpoint=False)

    def pt3(i, j):
        a, b = u[i % n_major], v[j % n_minor]
        return np.array([(R + r * np.cos(b)) * np.cos(a), (R + r * np.cos(b)) * np.sin(a), r * np.sin(b)]) + np.array(center)
    tris = []
    for i in range(n_major):
        for j in range(n_minor):
            p00, p10, p11, p01 = pt3(i, j), pt3(i + 1, j), pt3(i + 1, j + 1), pt3(i, j + 1)
            tris.append([p00, p10, p11])
            tris.append([p00, p11, p01])
    return np.array(tris)


def l_extrusion_triangles(height=0.5):
    """L-shaped prism: polygon (0,0)-(2,0)-(2,1)-(1,1)-(1,2)-(0,2) extruded along z in [0, height]; exact planar faces."""
    # caps: a fan over the ring (0,0) (2,0) (2,1) (1,1) (1,2) (0,2) (0,1) -- every cap edge is a ring edge or shared twice
    caps = [[(0, 0), (2, 0), (2, 1)], [(0, 0), (2, 1), (1, 1)], [(0, 0), (1, 1), (0, 1)], [(0, 1), (1, 1), (1, 2)],
            [(0, 1), (1, 2), (0, 2)]]
    tris = []
    for z, flip in ((0.0, True), (height, False)):
        for t in caps:
            pts = [np.array([x, y, z], dtype=np.float64) for x, y in (t[::-1] if flip else t)]
            tris.append(pts)
    # walls over the same ring, so that the surface is watertight
    ring = [(0, 0), (2, 0), (2, 1), (1, 1), (1, 2), (0, 2), (0, 1)]
    for i in range(len(ring)):
        (x0, y0), (x1, y1) = ring[i], ring[(i + 1) % len(ring)]
        a, b = np.array([x0, y0, 0.0]), np.array([x1, y1, 0.0])
        c, d = np.array([x1, y1, height]), np.array([x0, y0, height])
        tris.append([a, b, c])
        tris.append([a, c, d])
    return np.array(tris)


def in_l_extrusion(p, height=0.5, margin=0.0):
    """Analytic classification: +1 inside by more than `margin`, -1 outside by more than `margin`, 0 in the band."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]

    def box(x0, x1, y0, y1, m):
        return (x > x0 + m) & (x < x1 - m) & (y > y0 + m) & (y < y1 - m) & (z > m) & (z < height - m)
    inside = box(0, 2, 0, 1, margin) | box(0, 1, 0, 2, margin)
    grown = box(0, 2, 0, 1, -margin) | box(0, 1, 0, 2, -margin)
    return np.where(inside, 1, np.where(~grown, -1, 0))
