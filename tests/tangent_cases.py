"""Triangle soups (three corners per face: position, normal, uv) that exercise every rule of the tangent-space generator.
Shared by tests/test_tangent_space.py and tests/golden/make_tangent_golden.py."""
import numpy as np


def _soup(pos, nrm, uv, idx):
    idx = np.asarray(idx).reshape(-1)
    return (np.ascontiguousarray(pos[idx], np.float32), np.ascontiguousarray(nrm[idx], np.float32),
            np.ascontiguousarray(uv[idx], np.float32))


def sphere(stacks=12, slices=16, mirror=False):
    """Smooth uv sphere: shared vertices, a uv seam, poles with collapsed uv edges; `mirror` flips u on one half."""
    th = np.linspace(0, np.pi, stacks + 1)
    ph = np.linspace(0, 2 * np.pi, slices + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    pos = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], -1).reshape(-1, 3)
    uv = np.stack([P / (2 * np.pi), T / np.pi], -1).reshape(-1, 2)
    if mirror:
        uv[:, 0] = np.where(uv[:, 0] > 0.5, 1.0 - uv[:, 0], uv[:, 0])
    idx = []
    for i in range(stacks):
        for j in range(slices):
            a, b = i * (slices + 1) + j, (i + 1) * (slices + 1) + j
            idx += [a, b, a + 1, a + 1, b, b + 1]
    return _soup(pos, pos.copy(), uv, idx)


def grid(n=9, seed=0, flat=False):
    """Wavy height field with shared smooth normals (or flat per-face normals), uv rotated and sheared."""
    rs = np.random.default_rng(seed)
    x, z = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n), indexing="ij")
    y = 0.3 * np.sin(3 * x) * np.cos(2 * z) + 0.05 * rs.normal(size=x.shape)
    pos = np.stack([x, y, z], -1).reshape(-1, 3)
    gy_x, gy_z = np.gradient(y, 2 / (n - 1))
    nrm = np.stack([-gy_x, np.ones_like(y), -gy_z], -1).reshape(-1, 3)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = np.stack([0.7 * x + 0.4 * z, -0.2 * x + 0.9 * z], -1).reshape(-1, 2)
    idx = []
    for i in range(n - 1):
        for j in range(n - 1):
            a = i * n + j
            idx += [a, a + 1, a + n, a + 1, a + n + 1, a + n]
    p, nn, t = _soup(pos, nrm, uv, idx)
    if flat:
        f = p.reshape(-1, 3, 3)
        fn = np.cross(f[:, 1] - f[:, 0], f[:, 2] - f[:, 0])
        fn /= np.linalg.norm(fn, axis=1, keepdims=True)
        nn = np.repeat(fn, 3, axis=0).astype(np.float32)
    return p, nn, t


def hostile(seed=1):
    """The grid with the special cases planted: coincident positions, zero uv area, a constant-uv triangle, mirrored uv
    islands next to regular ones, a fin of three triangles on one edge, and a face listed twice."""
    p, n, t = grid(8, seed)
    p, n, t = p.copy(), n.copy(), t.copy()
    f = lambda k: slice(3 * k, 3 * k + 3)  # noqa: E731
    p[3 * 5 + 1] = p[3 * 5]                                   # face 5: two coincident positions
    p[f(17)] = p[3 * 17]                                      # face 17: a point
    t[3 * 9 + 2] = 0.5 * (t[3 * 9] + t[3 * 9 + 1])            # face 9: collinear uv
    t[f(12)] = t[3 * 12]                                      # face 12: constant uv
    for k in (20, 21, 22, 23, 40, 41):                        # mirrored islands
        t[f(k), 0] = -t[f(k), 0]
    # a fin: one more triangle standing on the first edge of face 30, and face 31 repeated
    a, b = p[3 * 30], p[3 * 30 + 1]
    fin_p = np.stack([b, a, 0.5 * (a + b) + [0, 0.4, 0]]).astype(np.float32)
    fin_n = np.stack([n[3 * 30 + 1], n[3 * 30], [1.0, 0.0, 0.0]]).astype(np.float32)
    fin_t = np.stack([t[3 * 30 + 1], t[3 * 30], t[3 * 30] + [0.0, 0.3]]).astype(np.float32)
    p = np.concatenate([p, fin_p, p[f(31)]])
    n = np.concatenate([n, fin_n, n[f(31)]])
    t = np.concatenate([t, fin_t, t[f(31)]])
    return np.ascontiguousarray(p, np.float32), np.ascontiguousarray(n, np.float32), np.ascontiguousarray(t, np.float32)


def random_soup(faces=300, seed=2):
    """Unrelated random triangles that share a small pool of vertices: plenty of fins, inconsistent windings, open edges."""
    rs = np.random.default_rng(seed)
    pool = 60
    pos = rs.uniform(-1, 1, (pool, 3))
    nrm = rs.normal(size=(pool, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = rs.uniform(0, 1, (pool, 2))
    idx = rs.integers(0, pool, (faces, 3))
    return _soup(pos, nrm, uv, idx)


CASES = {
    "sphere": lambda: sphere(),
    "sphere_mirrored": lambda: sphere(10, 14, True),
    "grid_smooth": lambda: grid(9, 0),
    "grid_flat": lambda: grid(9, 3, True),
    "hostile": lambda: hostile(),
    "random_soup": lambda: random_soup(),
}
