"""The rasterising producers: G-buffer pass and omni shadow cubes.

CPU tests pin the restatement's rules (oracle/althea_oracle_raster.cpp) on cases with known answers: pixel-aligned quads
(top-left rule), watertight shared edges, draw-order depth ties, culling, triangles through the eye plane, and the consistency
of the produced shadow cubes with the cube lookup the deferred pass performs. GPU tests compare the CUDA path with the
restatement: depth bit for bit, shaded attributes within tolerance.
"""
import os

import numpy as np
import pytest

from althea_b200 import model, scene
from helpers import REFERENCE

NONE = 0xFFFFFFFF


def _camera(W, H, pos=(0.0, 0.0, 4.0), yaw=0.0, pitch=0.0):
    g = scene.make_uniforms(W, H, pos=pos, yaw=yaw, pitch=pitch)
    return g, list(g.projection), list(g.view)


def _tri_prim(points, uv=None, front_cw=False):
    p = np.asarray(points, np.float32).reshape(-1, 3)
    n = np.tile(np.array([[0.0, 0.0, 1.0]], np.float32), (len(p), 1))
    v = model.make_vertices(p, n, uv)
    return model.PrimitiveData(v, np.arange(len(p), dtype=np.uint32), front_face_clockwise=front_cw)


def _ortho_like(W, H):
    """A camera whose pixel grid maps to world units exactly: eye at z = d, plane z = 0 spans W x H world units."""
    # perspective with fov such that at distance d the view height is H: tan(fov/2) = H / (2 d)
    d = 64.0
    fov = 2.0 * np.degrees(np.arctan(H / (2.0 * d)))
    g = scene.make_uniforms(W, H, pos=(0.0, 0.0, d), yaw=0.0, pitch=0.0, fov_deg=float(fov))
    return g, list(g.projection), list(g.view)


def _world_of_pixel_corner(W, H, x, y):
    """world position on the plane z = 0 of pixel-grid coordinate (x, y) (corner coordinates, y down)."""
    return (x - W / 2.0, H / 2.0 - y, 0.0)


# ---- CPU: the restatement's rules ------------------------------------------------------------------------------------------
def test_top_left_rule_on_a_pixel_aligned_quad(oracle):
    W, H = 32, 16
    g, proj, view = _ortho_like(W, H)
    c = [_world_of_pixel_corner(W, H, x, y) for x, y in ((4, 3), (4, 9), (12, 9), (12, 3))]  # counter-clockwise on screen (y down)
    q = model.quad(c)
    out = oracle.draw_gbuffer(proj, view, [q], W, H)
    cov = out["tri"] != NONE
    expect = np.zeros((H, W), bool)
    expect[3:9, 4:12] = True  # pixel centres x + 0.5 in [4, 12), y + 0.5 in [3, 9)
    assert np.array_equal(cov, expect)


def test_shared_edges_are_watertight_and_exclusive(oracle):
    """A fan of triangles around an interior vertex at an awkward position: every pixel inside the outline is covered, and
    each by the triangle whose interior holds its centre (no double hits are observable through ids, so: no holes)."""
    W, H = 64, 48
    g, proj, view = _camera(W, H, pos=(0.1, -0.07, 3.0))
    rng = np.random.default_rng(3)
    centre = np.array([0.013, 0.021, 0.0])
    ring = [np.array([np.cos(a), np.sin(a), 0.0]) * (1.0 + 0.2 * rng.random()) for a in np.linspace(0, 2 * np.pi, 13)[:-1]]
    pts = []
    for i in range(12):
        pts += [centre, ring[i], ring[(i + 1) % 12]]
    out = oracle.draw_gbuffer(proj, view, [_tri_prim(pts)], W, H)
    cov = out["tri"] != NONE
    # the union is a star-shaped polygon: compare with a point-in-polygon test away from the outline
    ys, xs = np.nonzero(cov)
    assert cov.sum() > 400
    # no holes: every covered row is one contiguous run
    for y in np.unique(ys):
        row = np.nonzero(cov[y])[0]
        assert row[-1] - row[0] + 1 == len(row)


def test_depth_ties_keep_the_first_drawn(oracle):
    W, H = 32, 32
    g, proj, view = _camera(W, H)
    tri = [(-1, -1, 0), (1, -1, 0), (0, 1, 0)]
    a, b = _tri_prim(tri), _tri_prim(tri)
    out = oracle.draw_gbuffer(proj, view, [a, b], W, H)
    ids = out["tri"][out["tri"] != NONE]
    assert len(ids) > 50 and (ids == 0).all()
    # a nearer triangle drawn later wins
    c = _tri_prim([(-1, -1, 0.5), (1, -1, 0.5), (0, 1, 0.5)])
    out = oracle.draw_gbuffer(proj, view, [a, c], W, H)
    assert (out["tri"][H // 2, W // 2] == 1)


def test_back_faces_are_culled_unless_the_front_face_is_clockwise(oracle):
    W, H = 32, 32
    g, proj, view = _camera(W, H)
    ccw = [(-1, -1, 0), (1, -1, 0), (0, 1, 0)]  # counter-clockwise seen from +z (the camera side)
    cw = [ccw[0], ccw[2], ccw[1]]
    assert (oracle.draw_gbuffer(proj, view, [_tri_prim(ccw)], W, H)["tri"] != NONE).sum() > 50
    assert (oracle.draw_gbuffer(proj, view, [_tri_prim(cw)], W, H)["tri"] != NONE).sum() == 0
    assert (oracle.draw_gbuffer(proj, view, [_tri_prim(cw, front_cw=True)], W, H)["tri"] != NONE).sum() > 50
    assert (oracle.draw_gbuffer(proj, view, [_tri_prim(ccw, front_cw=True)], W, H)["tri"] != NONE).sum() == 0


def test_triangle_through_the_eye_plane(oracle):
    """A ground quad that extends behind the camera: the visible part is everything below the horizon, depth grows towards it,
    reconstructed positions lie on the plane."""
    W, H = 64, 36
    g, proj, view = _camera(W, H, pos=(0.0, 1.0, 0.0))
    floor = model.quad([(-50, 0, 50), (50, 0, 50), (50, 0, -50), (-50, 0, -50)])
    out = oracle.draw_gbuffer(proj, view, [floor], W, H)
    cov = out["tri"] != NONE
    assert cov[H // 2 + 1:].all() and not cov[: H // 2 - 1].any()
    assert np.abs(out["position"][cov][:, 1]).max() < 1e-3
    col = out["depth"][H // 2 + 1:, W // 2]
    assert (np.diff(col) < 0).all()  # nearer (smaller depth) towards the bottom of the screen
    n = out["normal"][cov][:, :3]
    assert np.abs(n - np.array([0, 1, 0])).max() < 5e-3  # normal1x1.png is (128, 128, 255): a 0.0039 tilt


def test_material_fetch(oracle):
    """Flat normal map and white defaults; factors, the .bg swizzle of metallic-roughness, ao = 0, alpha in normal.a / mro.a."""
    W, H = 16, 16
    g, proj, view = _camera(W, H, pos=(0.0, 0.0, 1.0))
    q = model.quad([(-2, -2, 0), (2, -2, 0), (2, 2, 0), (-2, 2, 0)])
    q.material = model.MaterialData(baseColorFactor=(0.5, 0.25, 1.0, 0.8), metallicFactor=0.3, roughnessFactor=0.6)
    mr = np.zeros((4, 4, 4), np.uint8)
    mr[...] = (10, 204, 102, 255)  # roughness from .g = 0.8, metallic from .b = 0.4
    q.material.metallicRoughnessTexture = model.TextureData.from_rgba8(mr, model.sampler_word())
    out = oracle.draw_gbuffer(proj, view, [q], W, H)
    assert (out["tri"] != NONE).all()
    # every pipeline of the reference blends its colour attachments (SRC_ALPHA, ONE_MINUS_SRC_ALPHA; alpha: ONE, ZERO,
    # Src/GraphicsPipeline.cpp:138-154): over the 0 clear a fragment with alpha 0.8 leaves rgb * 0.8, alpha 0.8
    assert np.array_equal(out["albedo"][8, 8], [round(0.5 * 0.8 * 255), round(0.25 * 0.8 * 255), round(0.8 * 255), 204])
    assert np.array_equal(out["mro"][8, 8], [round(0.4 * 0.3 * 0.8 * 255), round(0.8 * 0.6 * 0.8 * 255), 0, 204])
    assert np.abs(out["normal"][8, 8] - [0, 0, 0.8, 0.8]).max() < 5e-3  # normal1x1.png is (128, 128, 255): a 0.0039 tilt
    # alpha below the cutoff discards every fragment
    q.material.baseColorFactor = (1, 1, 1, 0.4)
    assert (oracle.draw_gbuffer(proj, view, [q], W, H)["tri"] == NONE).all()


def test_restatement_against_a_float64_screen_space_rasteriser(oracle):
    """The restatement works with edge functions in homogeneous clip space (fp32). An independent float64 rasteriser written the
    textbook way (project the vertices, 2-D edge functions on the screen, perspective-correct interpolation through 1/w)
    must agree with it on which triangle owns every pixel whose centre is not within 1e-6 of an edge, on depth to 1e-6 and on
    the interpolated world position to 1e-4. Random opaque triangles in front of the camera."""
    rng = np.random.default_rng(41)
    W, H = 80, 60
    g, proj, view = _camera(W, H, pos=(0.2, -0.1, 5.0), yaw=0.07, pitch=-0.04)
    tris = rng.uniform(-2.5, 2.5, (60, 3, 3))
    tris[:, :, 2] = rng.uniform(-2.0, 2.0, (60, 3))
    prim = _tri_prim(tris.reshape(-1, 3).astype(np.float32))
    out = oracle.draw_gbuffer(proj, view, [prim], W, H)
    P = np.array(proj, np.float64).reshape(4, 4).T
    V = np.array(view, np.float64).reshape(4, 4).T
    pts = prim.vertices[:, :3].astype(np.float64).reshape(-1, 3, 3)
    clip = np.einsum("ij,tkj->tki", P @ V, np.concatenate([pts, np.ones(pts.shape[:2] + (1,))], -1))
    assert (clip[..., 3] > 0.1).all()
    ndc = clip[..., :3] / clip[..., 3:4]
    sx, sy = (ndc[..., 0] + 1) * 0.5 * W, (ndc[..., 1] + 1) * 0.5 * H
    ys, xs = np.mgrid[0:H, 0:W]
    cx, cy = xs + 0.5, ys + 0.5
    best_z = np.full((H, W), np.inf)
    best_t = np.full((H, W), -1)
    best_p = np.zeros((H, W, 3))
    near_edge = np.zeros((H, W), bool)
    for t in range(len(pts)):
        x0, x1, x2 = sx[t]
        y0, y1, y2 = sy[t]
        area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        if area >= 0:  # y-down framebuffer: positive = clockwise as seen = back face (the default front face is counter-clockwise)
            continue
        w0 = ((x1 - cx) * (y2 - cy) - (x2 - cx) * (y1 - cy)) / area
        w1 = ((x2 - cx) * (y0 - cy) - (x0 - cx) * (y2 - cy)) / area
        w2 = 1.0 - w0 - w1
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        near_edge |= (np.minimum(np.minimum(np.abs(w0), np.abs(w1)), np.abs(w2)) < 1e-6)
        z = w0 * ndc[t, 0, 2] + w1 * ndc[t, 1, 2] + w2 * ndc[t, 2, 2]  # z/w is affine on the screen
        iw = w0 / clip[t, 0, 3] + w1 / clip[t, 1, 3] + w2 / clip[t, 2, 3]
        b = np.stack([w0 / clip[t, 0, 3], w1 / clip[t, 1, 3], w2 / clip[t, 2, 3]], -1) / iw[..., None]
        pos = b @ pts[t]
        win = inside & (z < best_z) & (z >= 0) & (z <= 1)
        best_z[win], best_t[win], best_p[win] = z[win], t, pos[win]
    ok = ~near_edge
    mine = np.where(best_t >= 0, best_t, NONE).astype(np.uint64)
    theirs = out["tri"].astype(np.uint64)
    if (mine[ok] != theirs[ok]).mean() > 0.5:  # the other winding convention: redo is pointless, fail loudly
        raise AssertionError("front-face convention differs between the two rasterisers")
    assert (mine[ok] == theirs[ok]).mean() > 0.999
    same = ok & (mine == theirs) & (best_t >= 0)
    assert same.sum() > 500
    assert np.abs(out["depth"][same] - best_z[same]).max() < 2.5e-7  # a few ulp of a depth near 1 (6e-8 each)
    assert np.abs(out["position"][same][:, :3] - best_p[same]).max() < 1e-4


def test_translucent_fragments_blend_in_draw_order(oracle):
    """Alpha blending of the G-buffer attachments: a translucent quad drawn AFTER an opaque one behind it mixes with it; drawn
    BEFORE it, the opaque quad fails the depth test there and the translucent one stays mixed with the clear colour; two
    translucent layers compound, each step rounded to the attachment's format."""
    W, H = 16, 16
    g, proj, view = _camera(W, H, pos=(0.0, 0.0, 2.0))
    back = model.quad([(-3, -3, -1), (3, -3, -1), (3, 3, -1), (-3, 3, -1)])
    back.material = model.MaterialData(baseColorFactor=(1.0, 0.0, 0.0, 1.0), metallicFactor=1.0, roughnessFactor=1.0)
    front = model.quad([(-3, -3, 0), (3, -3, 0), (3, 3, 0), (-3, 3, 0)])
    front.material = model.MaterialData(baseColorFactor=(0.0, 1.0, 0.0, 0.6), metallicFactor=0.0, roughnessFactor=0.5)
    a = oracle.draw_gbuffer(proj, view, [back, front], W, H)
    assert np.array_equal(a["albedo"][8, 8], [round(0.4 * 255), round(0.6 * 255), 0, 153])
    assert np.array_equal(a["mro"][8, 8], [round(0.4 * 255), round((0.5 * 0.6 + 0.4) * 255), 0, 153])
    assert (a["tri"][8, 8] >= 2) and a["depth"][8, 8] < oracle.draw_gbuffer(proj, view, [back], W, H)["depth"][8, 8]
    b = oracle.draw_gbuffer(proj, view, [front, back], W, H)
    assert np.array_equal(b["albedo"][8, 8], [0, round(0.6 * 255), 0, 153])
    mid = model.quad([(-3, -3, -0.5), (3, -3, -0.5), (3, 3, -0.5), (-3, 3, -0.5)])
    mid.material = model.MaterialData(baseColorFactor=(0.0, 0.0, 1.0, 0.5))
    c = oracle.draw_gbuffer(proj, view, [back, mid, front], W, H)
    under = np.array([round(0.5 * 255), 0, round(0.5 * 255)]) / 255.0  # back under mid, stored as UNORM8
    want = [round(float(under[0]) * 0.4 * 255), round((0.6 + float(under[1]) * 0.4) * 255), round(float(under[2]) * 0.4 * 255), 153]
    assert np.abs(c["albedo"][8, 8].astype(int) - np.array(want)).max() <= 1


def test_mip_selection_and_wrap(oracle):
    """A 2-texel checker minified heavily averages to grey through the mip chain; magnified it keeps its two colours."""
    W, H = 32, 32
    tex = model.checker_texture(size=64, cells=64, a=(255, 255, 255, 255), b=(0, 0, 0, 255), sampler=model.sampler_word(srgb=False))
    far = model.quad([(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)], uv_scale=8.0)
    far.material.baseTexture = tex
    g, proj, view = _camera(W, H, pos=(0.0, 0.0, 6.0))
    out = oracle.draw_gbuffer(proj, view, [far], W, H)
    a = out["albedo"][out["tri"] != NONE][:, 0].astype(int)
    assert len(a) > 20 and np.abs(a - 128).max() <= 2
    near = model.quad([(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)], uv_scale=4.0 / 64.0)
    near.material.baseTexture = tex
    g, proj, view = _camera(W, H, pos=(0.0, 0.0, 1.2))
    out = oracle.draw_gbuffer(proj, view, [near], W, H)
    a = out["albedo"][..., 0]
    assert a.max() >= 240 and a.min() <= 15


def test_wrap_modes_and_nearest_filtering(oracle):
    """uv running from -1 to 2 across a quad over a 4-texel ramp: REPEAT shows the ramp three times, CLAMP_TO_EDGE holds the end
    texels outside [0, 1], MIRRORED_REPEAT reverses the outer copies; NEAREST magnification returns texel values unblended."""
    W, H = 96, 8
    ramp = np.zeros((1, 4, 4), np.uint8)
    ramp[0, :, 0] = (0, 85, 170, 255)
    ramp[..., 3] = 255
    ramp = np.repeat(ramp, 4, axis=0)
    g, proj, view = _ortho_like(W, H)
    corners = [_world_of_pixel_corner(W, H, x, y) for x, y in ((0, 0), (0, H), (W, H), (W, 0))]

    def render(wrap):
        q = model.quad(corners)
        # quad(): uv (0,0) (1,0) (1,1) (0,1) at the corners in order; remap u to run -1 .. 2 along x (corner order: x = 0, 0, W, W)
        q.vertices[:, 12] = np.array([-1.0, -1.0, 2.0, 2.0], np.float32)
        q.vertices[:, 13] = np.array([0.0, 1.0, 1.0, 0.0], np.float32)
        q.material.baseTexture = model.TextureData.from_rgba8(ramp, model.sampler_word(wrap, wrap, mag_nearest=True, mip_mode=model.MIP_NONE, srgb=False))
        out = oracle.draw_gbuffer(proj, view, [q], W, H)
        assert (out["tri"] != NONE).all()
        return out["albedo"][H // 2, :, 0].astype(int)

    # 96 pixels over 3 uv units: 32 pixels per unit, 8 pixels per texel
    texel = np.array([0, 85, 170, 255])
    unit = np.repeat(texel, 8)
    assert np.array_equal(render(model.WRAP_REPEAT), np.tile(unit, 3))
    assert np.array_equal(render(model.WRAP_CLAMP), np.concatenate([np.full(32, 0), unit, np.full(32, 255)]))
    assert np.array_equal(render(model.WRAP_MIRROR), np.concatenate([unit[::-1], unit, unit[::-1]]))


def _shadow_scene():
    """A closed room (inward-facing walls) with a sphere in it."""
    sp = model.uv_sphere(0.8, (0.5, 0.6, -0.4), 10, 20)
    floor = model.quad([(-6, -1, 6), (6, -1, 6), (6, -1, -6), (-6, -1, -6)])
    ceil = model.quad([(-6, 5, -6), (6, 5, -6), (6, 5, 6), (-6, 5, 6)])
    back = model.quad([(-6, -1, -5), (6, -1, -5), (6, 6, -5), (-6, 6, -5)])
    front = model.quad([(6, -1, 6), (-6, -1, 6), (-6, 6, 6), (6, 6, 6)])
    left = model.quad([(-6, -1, 6), (-6, -1, -6), (-6, 6, -6), (-6, 6, 6)])
    right = model.quad([(6, -1, -6), (6, -1, 6), (6, 6, 6), (6, 6, -6)])
    return [sp, floor, ceil, back, front, left, right]


def _cube_lookup(cubes, d):
    """The Vulkan cube face selection and (s, t) the deferred pass uses (frame_kernels.cu sampleShadowCube), nearest texel."""
    ax, ay, az = np.abs(d)
    if ax >= ay and ax >= az:
        face, sc, tc, ma = (0, -d[2], -d[1], ax) if d[0] >= 0 else (1, d[2], -d[1], ax)
    elif ay >= az:
        face, sc, tc, ma = (2, d[0], d[2], ay) if d[1] >= 0 else (3, d[0], -d[2], ay)
    else:
        face, sc, tc, ma = (4, d[0], -d[1], az) if d[2] >= 0 else (5, -d[0], -d[1], az)
    res = cubes.shape[-1]
    s, t = 0.5 * sc / ma + 0.5, 0.5 * tc / ma + 0.5
    return face, min(int(s * res), res - 1), min(int(t * res), res - 1)


def test_shadow_cubes_agree_with_the_deferred_pass_lookup(oracle):
    """The producer's faces against the lookup the consumer performs (PointLights.glsl samples the cube with (L.x, -L.y, -L.z),
    L = light - surface): for points of a closed room, the texel the consumer reads must hold the point's own distance when
    the point is lit and a smaller one when the sphere shadows it. The +-X and +-Z faces satisfy this as rendered. The
    reference's +-Y face cameras are rotated by 180 degrees about the face axis (yaw = 180 at pitch = +-90,
    Src/PointLight.cpp:100-108): a faithful producer reproduces that, so on those faces the agreement holds after rotating the
    face, and NOT as rendered (asserted too: it documents the reference's behaviour)."""
    pc = model.point_light_constants()
    views = np.array([list(pc.views[f]) for f in range(6)], np.float32)
    light = np.array([0.3, 2.0, 0.7], np.float32)
    lights = np.zeros((1, 8), np.float32)
    lights[0, :3] = light
    res = 128
    cubes = oracle.draw_shadow_cubes(lights, list(pc.projection), views, _shadow_scene(), res)[0]
    assert all((cubes[f] < 1).mean() > 0.99 for f in range(6))  # a closed room: every face sees geometry
    centre, radius = np.array([0.5, 0.6, -0.4]), 0.8

    def occluded(p):
        d = light - p
        ln = np.linalg.norm(d)
        d = d / ln
        oc = p - centre
        b, c = np.dot(oc, d), np.dot(oc, oc) - radius * radius
        disc = b * b - c
        return disc > 0 and 0 < -b - np.sqrt(disc) < ln

    rng = np.random.default_rng(11)
    stats = {flip: {f: [0, 0, 0, 0] for f in range(6)} for flip in (False, True)}  # lit ok, lit n, shadowed ok, shadowed n
    for _ in range(8000):
        k = rng.integers(0, 6)
        a, b = rng.uniform(-5.5, 5.5), rng.uniform(-4.5, 5.5)
        h = rng.uniform(-0.5, 4.5)
        p = np.array([(a, -1.0, b), (a, 5.0, b), (a, h, -5.0), (a, h, 6.0), (-6.0, h, b), (6.0, h, b)][k])
        L = light - p
        dist = np.linalg.norm(L)
        face, sx, sy = _cube_lookup(cubes, np.array([L[0], -L[1], -L[2]]) / dist)
        shadowed = occluded(p)
        for flip in (False, True):
            x, y = (res - 1 - sx, res - 1 - sy) if (flip and face in (2, 3)) else (sx, sy)
            stored = cubes[face, y, x] * 1000.0
            st = stats[flip][face]
            if shadowed:
                st[3] += 1
                st[2] += stored < dist - 0.3
            else:
                st[1] += 1
                st[0] += abs(stored - dist) < 0.02 * dist + 0.05
    for f in range(6):
        lit_ok, lit_n, sh_ok, sh_n = stats[True][f]
        assert lit_n > 100 and lit_ok / lit_n > 0.99, (f, stats[True][f])
        if sh_n > 50:
            assert sh_ok / sh_n > 0.9, (f, stats[True][f])
    lit_ok, lit_n, sh_ok, sh_n = stats[False][3]  # the -Y face as rendered does not line up with the lookup
    assert sh_n > 50 and sh_ok / sh_n < 0.2 and lit_ok / lit_n < 0.9


@pytest.mark.skipif(not os.path.isfile(os.path.join(REFERENCE, "Content/Models/DamagedHelmet.glb")), reason="reference not mounted (GPU box)")
def test_glb_loader_on_the_reference_asset(oracle):
    prims = model.load_glb(os.path.join(REFERENCE, "Content/Models/DamagedHelmet.glb"), max_texture_size=256)
    assert sum(p.triangle_count for p in prims) == 46356 // 3  # 46 356 indices
    p = prims[0]
    assert p.material.baseTexture is not None and p.material.normalTexture is not None and p.material.metallicRoughnessTexture is not None
    nrm = p.vertices[:, 9:12]
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1).max() < 1e-3
    # the file ships no TANGENT: generated as the reference does (tests/test_tangent_space.py holds them to its MikkTSpace
    # build); all but a handful of corners on zero-uv-area triangles are unit and perpendicular to the normal
    t = p.vertices[:, 3:6]
    assert (np.abs(np.linalg.norm(t, axis=1) - 1) > 1e-5).sum() < 20   # the generator leaves a tangent it cannot derive at zero
    assert (np.abs(np.sum(t * nrm, axis=1)) > 1e-3).sum() < 20
    b = p.vertices[:, 6:9]
    assert np.abs(np.abs(np.sum(b * np.cross(nrm, t), axis=1)) - np.sum(np.cross(nrm, t) ** 2, axis=1)).max() < 1e-5
    W, H = 96, 54
    g, proj, view = _camera(W, H, pos=(0.0, 0.0, 3.0))
    out = oracle.draw_gbuffer(proj, view, prims, W, H)
    cov = out["tri"] != NONE
    assert 0.05 < cov.mean() < 0.6
    assert np.abs(np.linalg.norm(out["normal"][cov][:, :3], axis=1) - 1).max() < 2e-3  # as stored: RGBA16F


def _write_glb(path, pos, nrm, uv, idx, png_rgba, with_tangents=False):
    """A minimal binary glTF 2.0: one node (translated), one mesh primitive, one material with a PNG base colour texture."""
    import io
    import json
    import struct

    from PIL import Image as PILImage
    buf = io.BytesIO()
    PILImage.fromarray(png_rgba, "RGBA").save(buf, format="PNG")
    png = buf.getvalue()
    chunks, views, accessors = [], [], []

    def add(data, target=None):
        off = sum(len(c) for c in chunks)
        pad = (-len(data)) % 4
        chunks.append(data + b"\0" * pad)
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if target:
            v["target"] = target
        views.append(v)
        return len(views) - 1

    def acc(arr, ctype, typ, target=None):
        accessors.append({"bufferView": add(np.ascontiguousarray(arr).tobytes(), target), "componentType": ctype, "count": len(arr), "type": typ})
        return len(accessors) - 1

    a_pos, a_nrm, a_uv = acc(pos.astype(np.float32), 5126, "VEC3", 34962), acc(nrm.astype(np.float32), 5126, "VEC3", 34962), acc(uv.astype(np.float32), 5126, "VEC2", 34962)
    a_idx = acc(idx.astype(np.uint16), 5123, "SCALAR", 34963)
    attrs = {"POSITION": a_pos, "NORMAL": a_nrm, "TEXCOORD_0": a_uv}
    if with_tangents:
        tan = np.tile(np.array([[1.0, 0.0, 0.0, 1.0]], np.float32), (len(pos), 1))
        attrs["TANGENT"] = acc(tan, 5126, "VEC4", 34962)
    img_view = add(png)
    gltf = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
            "nodes": [{"mesh": 0, "translation": [0.5, 0.0, -1.0], "scale": [2.0, 2.0, 2.0]}],
            "meshes": [{"primitives": [{"attributes": attrs, "indices": a_idx, "material": 0}]}],
            "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "baseColorFactor": [1.0, 0.5, 0.25, 1.0], "metallicFactor": 0.2,
                                                     "roughnessFactor": 0.7}, "alphaCutoff": 0.3}],
            "textures": [{"source": 0, "sampler": 0}], "samplers": [{"wrapS": 33071, "wrapT": 33648, "minFilter": 9729, "magFilter": 9728}],
            "images": [{"bufferView": img_view, "mimeType": "image/png"}], "bufferViews": views, "accessors": accessors,
            "buffers": [{"byteLength": sum(len(c) for c in chunks)}]}
    js = json.dumps(gltf).encode()
    js += b" " * ((-len(js)) % 4)
    binc = b"".join(chunks)
    with open(path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(binc)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(binc), 0x004E4942) + binc)


@pytest.mark.parametrize("with_tangents", [False, True])
def test_glb_loader_on_a_synthetic_file(tmp_path, oracle, with_tangents):
    """load_glb on a file written here (the reference's assets are not on the GPU box): node transform, material factors, sampler
    translation, texture decode + mip chain, and the two vertex paths of Primitive.cpp (tangents from the file: indexed vertices;
    no tangents: vertices de-indexed, tangents generated)."""
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = np.array([[0, 1], [1, 1], [1, 0], [0, 0]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    tex = np.zeros((8, 8, 4), np.uint8)
    tex[..., 0], tex[..., 3] = 200, 255
    tex[:4, :, 1] = 255
    path = str(tmp_path / "quad.glb")
    _write_glb(path, pos, nrm, uv, idx, tex, with_tangents)
    prims = model.load_glb(path)
    assert len(prims) == 1
    p = prims[0]
    assert p.triangle_count == 2
    assert len(p.vertices) == (4 if with_tangents else 6)  # Primitive.cpp:147: generated tangents need unshared vertices
    m = p.material
    assert m.baseColorFactor == (1.0, 0.5, 0.25, 1.0) and m.metallicFactor == pytest.approx(0.2) and m.roughnessFactor == pytest.approx(0.7)
    assert m.alphaCutoff == pytest.approx(0.3) and m.normalTexture is None and m.metallicRoughnessTexture is None
    t = m.baseTexture
    assert t.width == 8 and t.height == 8 and len(t.levels) == 1  # minFilter LINEAR: no mips (Src/Sampler.cpp:78-85)
    assert t.sampler == model.sampler_word(model.WRAP_CLAMP, model.WRAP_MIRROR, True, False, model.MIP_NONE, True)
    assert np.array_equal(t.levels[0], tex)
    assert np.allclose(p.model, np.array([[2, 0, 0, 0.5], [0, 2, 0, 0], [0, 0, 2, -1], [0, 0, 0, 1]], np.float32))
    tang, bit, n = p.vertices[:, 3:6], p.vertices[:, 6:9], p.vertices[:, 9:12]
    assert np.allclose(n, [0, 0, 1]) and np.allclose(np.abs(tang), [1, 0, 0], atol=1e-6) and np.allclose(np.abs(bit), [0, 1, 0], atol=1e-6)
    # and it renders: the quad (scaled by 2, moved) in front of the camera, coloured by factor * texture
    W, H = 48, 32
    g, proj, view = _camera(W, H, pos=(0.5, 0.0, 3.0))
    out = oracle.draw_gbuffer(proj, view, prims, W, H)
    cov = out["tri"] != NONE
    assert 0.2 < cov.mean() < 0.9
    r = out["albedo"][cov][:, 0]
    assert r.min() > 100  # sRGB 200 -> linear 0.578 -> * 1.0 -> 147
    assert set(np.unique(out["albedo"][cov][:, 1] > 60)) == {False, True}  # the green half of the texture and the other half


def test_gltf_json_container_loads_like_the_glb(tmp_path):
    """The same asset as a .gltf: JSON next to an external .bin, one image as a file (with a space in its name, percent-encoded in
    the uri) and the buffer optionally as a base64 data: URI. Same primitives as the binary container."""
    import base64
    import io
    import json
    import struct

    from PIL import Image as PILImage
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = np.array([[0, 1], [1, 1], [1, 0], [0, 0]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    tex = np.zeros((8, 8, 4), np.uint8)
    tex[..., 0], tex[..., 3] = 200, 255
    tex[:4, :, 1] = 255
    glb = str(tmp_path / "quad.glb")
    _write_glb(glb, pos, nrm, uv, idx, tex, True)
    want = model.load_glb(glb)[0]
    data = open(glb, "rb").read()
    jlen = struct.unpack_from("<I", data, 12)[0]
    gltf = json.loads(data[20:20 + jlen])
    binc = data[20 + jlen + 8:]
    img_view = gltf["bufferViews"][gltf["images"][0]["bufferView"]]
    png = binc[img_view["byteOffset"]: img_view["byteOffset"] + img_view["byteLength"]]
    (tmp_path / "sub").mkdir()
    open(str(tmp_path / "sub" / "base colour.png"), "wb").write(png)
    gltf["images"] = [{"uri": "sub/base%20colour.png"}]
    for variant, uri in (("file", "quad.bin"), ("data", "data:application/octet-stream;base64," + base64.b64encode(binc).decode())):
        gltf["buffers"] = [{"byteLength": len(binc), "uri": uri}]
        open(str(tmp_path / "quad.bin"), "wb").write(binc)
        path = str(tmp_path / ("quad_%s.gltf" % variant))
        open(path, "w").write(json.dumps(gltf))
        got = model.load_gltf(path)
        assert len(got) == 1
        assert np.array_equal(got[0].vertices, want.vertices) and np.array_equal(got[0].indices, want.indices)
        assert np.array_equal(got[0].model, want.model)
        assert np.array_equal(got[0].material.baseTexture.levels[0], want.material.baseTexture.levels[0])
        assert got[0].material.baseTexture.sampler == want.material.baseTexture.sampler
    open(str(tmp_path / "junk.gltf"), "wb").write(b"\x00\x01 not json")
    with pytest.raises(ValueError):
        model.load_gltf(str(tmp_path / "junk.gltf"))
    open(str(tmp_path / "other.gltf"), "w").write("{}")
    with pytest.raises(ValueError):
        model.load_gltf(str(tmp_path / "other.gltf"))


def test_skinned_primitive_is_transformed_like_the_vertex_shader(tmp_path, oracle):
    """Gltf.vert:37-54 on the host: a strip skinned to two joints (u8 joints, normalised u16 weights, indexed, with tangents).
    The loaded vertices must be sum_i w_i * (global(joint_i) * inverseBind(joint_i)) * position, evaluated here in float64 from the
    file's numbers, the primitive's own transform must be the identity (the mesh node's transform does not apply to a skinned
    primitive), and the result renders where that puts it."""
    import json
    import struct
    pos = np.array([[-0.5, 0, 0], [0.5, 0, 0], [-0.5, 1, 0], [0.5, 1, 0], [-0.5, 2, 0], [0.5, 2, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (6, 1))
    tan = np.tile(np.array([[1, 0, 0, 1]], np.float32), (6, 1))
    uv = pos[:, :2].copy()
    idx = np.array([0, 1, 2, 2, 1, 3, 2, 3, 4, 4, 3, 5], np.uint16)
    joints = np.array([[0, 1, 0, 0]] * 6, np.uint8)
    w1 = np.array([0.0, 0.0, 0.5, 0.5, 1.0, 1.0])
    weights = np.zeros((6, 4), np.uint16)
    weights[:, 1] = np.round(w1 * 65535)
    weights[:, 0] = 65535 - weights[:, 1]
    c, s_ = np.cos(0.6), np.sin(0.6)
    ibm = np.stack([np.eye(4), np.array([[1, 0, 0, 0], [0, 1, 0, -1.0], [0, 0, 1, 0], [0, 0, 0, 1]])]).astype(np.float32)
    chunks, views, accessors = [], [], []

    def acc(arr, ctype, typ, normalized=False):
        data = np.ascontiguousarray(arr).tobytes()
        off = sum(len(x) for x in chunks)
        chunks.append(data + b"\0" * ((-len(data)) % 4))
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(data)})
        a = {"bufferView": len(views) - 1, "componentType": ctype, "count": len(arr), "type": typ}
        if normalized:
            a["normalized"] = True
        accessors.append(a)
        return len(accessors) - 1

    attrs = {"POSITION": acc(pos, 5126, "VEC3"), "NORMAL": acc(nrm, 5126, "VEC3"), "TANGENT": acc(tan, 5126, "VEC4"),
             "TEXCOORD_0": acc(uv, 5126, "VEC2"), "JOINTS_0": acc(joints, 5121, "VEC4"), "WEIGHTS_0": acc(weights, 5123, "VEC4", True)}
    a_idx = acc(idx, 5123, "SCALAR")
    a_ibm = acc(np.stack([m.T for m in ibm]).reshape(2, 16), 5126, "MAT4")  # column-major
    quat = [0.0, 0.0, float(np.sin(0.3)), float(np.cos(0.3))]              # 0.6 rad about z
    gltf = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 1]}],
            "nodes": [{"mesh": 0, "skin": 0, "translation": [100.0, 0.0, 0.0]},           # must NOT move the skinned primitive
                      {"children": [2], "translation": [0.25, 0.0, -1.0]}, {"translation": [0.0, 1.0, 0.0], "rotation": quat}],
            "skins": [{"joints": [1, 2], "inverseBindMatrices": a_ibm}],
            "meshes": [{"primitives": [{"attributes": attrs, "indices": a_idx}]}],
            "bufferViews": views, "accessors": accessors, "buffers": [{"byteLength": sum(len(x) for x in chunks)}]}
    js = json.dumps(gltf).encode()
    js += b" " * ((-len(js)) % 4)
    binc = b"".join(chunks)
    path = str(tmp_path / "skinned.glb")
    with open(path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(binc)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(binc), 0x004E4942) + binc)
    prims = model.load_gltf(path)
    assert len(prims) == 1 and np.array_equal(prims[0].model, np.eye(4, dtype=np.float32))
    v = prims[0].vertices
    assert len(v) == 6 and np.array_equal(prims[0].indices, idx)
    g0 = np.eye(4)
    g0[:3, 3] = [0.25, 0.0, -1.0]
    l1 = np.array([[c, -s_, 0, 0], [s_, c, 0, 1.0], [0, 0, 1, 0], [0, 0, 0, 1]])
    m = [g0 @ ibm[0].astype(np.float64), g0 @ l1 @ ibm[1].astype(np.float64)]
    wn = weights.astype(np.float64) / 65535.0
    for k in range(6):
        blend = wn[k, 0] * m[0] + wn[k, 1] * m[1]
        want = blend @ np.append(pos[k], 1.0)
        assert np.abs(v[k, 0:3] - want[:3]).max() < 1e-6
        assert np.abs(v[k, 3:6] - blend[:3, :3] @ [1, 0, 0]).max() < 1e-6       # tangent
        assert np.abs(v[k, 9:12] - blend[:3, :3] @ [0, 0, 1]).max() < 1e-6      # normal
    assert np.allclose(v[:, 20:24], wn, atol=1e-7) and np.array_equal(v[:, 24:26].copy().view(np.uint16).reshape(6, 4), joints)
    # the top of the strip swings 0.6 rad about the second joint; the mesh node's x = 100 is nowhere to be seen
    assert v[:, 0].max() < 2.0 and v[4, 0] < v[0, 0]
    W, H = 64, 48
    g, proj, view = _camera(W, H, pos=(0.0, 1.0, 3.0))
    out = oracle.draw_gbuffer(proj, view, prims, W, H)
    cov = out["tri"] != NONE
    assert 0.02 < cov.mean() < 0.5
    assert np.abs(out["position"][cov][:, 2] + 1.0).max() < 1e-4                # the whole strip lies in the plane z = -1


_REFERENCE_MODELS = {  # primitives, triangles, distinct textures; SURVEY.md 8(f): Sponza = 103 primitives / 262 k triangles / 69 images
    "DamagedHelmet.glb": (1, 15452, 3), "AntiqueCamera.glb": (2, 20066, 6), "Buggy.glb": (236, 531955, 0),
    "MetalRoughSpheres.glb": (5, 501776, 2), "FlightHelmet/FlightHelmet.gltf": (6, 94722, 15), "Sponza/glTF/Sponza.gltf": (103, 262267, 69)}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "Content/Models")), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("name", sorted(_REFERENCE_MODELS))
def test_loader_reads_every_model_the_reference_ships(name):
    prims = model.load_gltf(os.path.join(REFERENCE, "Content/Models", name), max_texture_size=32)
    n_prims, n_tris, n_tex = _REFERENCE_MODELS[name]
    assert len(prims) == n_prims and sum(p.triangle_count for p in prims) == n_tris
    textures = {id(t) for p in prims for t in (p.material.baseTexture, p.material.normalTexture, p.material.metallicRoughnessTexture) if t is not None}
    assert len(textures) == n_tex
    for p in prims:
        v = p.vertices
        assert np.isfinite(v[:, :12]).all() and int(p.indices.max()) < len(v) and len(p.indices) % 3 == 0
        n = np.linalg.norm(v[:, 9:12], axis=1)
        assert (np.abs(n - 1) < 1e-2).mean() > 0.99            # unit normals (a few degenerate faces give NaN-free zeros)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REFERENCE, "Content/Models/Sponza/glTF/Sponza.gltf")), reason="reference not mounted (GPU box)")
def test_sponza_frame_through_the_restatement(oracle):
    """BASELINE configs[1]'s scene (Sponza) from its .gltf: a view down the atrium fills the frame, with cut-out foliage and
    several hundred distinct triangles' worth of detail; then the deferred chain runs on that G-buffer."""
    prims = model.load_gltf(os.path.join(REFERENCE, "Content/Models/Sponza/glTF/Sponza.gltf"), max_texture_size=32)
    W, H = 128, 72
    g = scene.make_uniforms(W, H, pos=(-8.0, 2.0, 0.0), yaw=-np.pi / 2, pitch=0.0)
    out = oracle.draw_gbuffer(list(g.projection), list(g.view), prims, W, H)
    cov = out["tri"] != NONE
    assert cov.mean() > 0.99
    assert len(np.unique(out["tri"][cov])) > 1000
    assert 2.0 < np.linalg.norm(out["position"][cov][:, :3] - [-8.0, 2.0, 0.0], axis=1).max() < 40.0   # the far end of the nave
    nrm = out["normal"][cov][:, :3]
    unit = np.abs(np.linalg.norm(nrm, axis=1) - out["normal"][cov][:, 3]) < 5e-3   # unit normal x coverage alpha, RGBA16F
    assert unit.mean() > 0.97                                   # the rest: foliage texels blended over what is behind them


def test_triangles_sharing_an_edge_never_both_claim_a_pixel(oracle):
    """Random quads split along a diagonal, random cameras: the two triangles' coverages are disjoint and their union is the
    coverage of drawing both (top-left rule + exactly opposite edge values on the shared edge)."""
    rng = np.random.default_rng(23)
    W, H = 72, 56
    for trial in range(12):
        c = rng.uniform(-1.5, 1.5, (4, 3))
        c[:, 2] = rng.uniform(-1.0, 1.0)  # planar in z up to a shear below
        c[:, 2] += 0.3 * c[:, 0]
        g, proj, view = _camera(W, H, pos=tuple(rng.uniform(-0.3, 0.3, 2)) + (4.0,), yaw=float(rng.uniform(-0.1, 0.1)), pitch=float(rng.uniform(-0.1, 0.1)))
        a = _tri_prim([c[0], c[1], c[2]])
        b = _tri_prim([c[0], c[2], c[3]])
        ca = oracle.draw_gbuffer(proj, view, [a], W, H)["tri"] != NONE
        cb = oracle.draw_gbuffer(proj, view, [b], W, H)["tri"] != NONE
        both = oracle.draw_gbuffer(proj, view, [a, b], W, H)["tri"] != NONE
        if ca.sum() == 0 or cb.sum() == 0:  # the quad faces away from this camera (or is bow-tied): nothing to check
            continue
        assert not (ca & cb).any(), trial
        assert np.array_equal(ca | cb, both), trial


# ---- GPU: the CUDA path against the restatement ---------------------------------------------------------------------------
def _gpu_gbuffer(ctx, uniforms, prims, W, H):
    import torch

    from althea_b200 import engine
    up = model.UploadedModel(ctx, prims)
    gb = engine.GBufferResources(ctx, W, H)
    engine.SceneToGBufferPass(ctx).draw(uniforms, up, gb)
    torch.cuda.synchronize()
    return {
        "depth": gb.depth.tensor.view(torch.float32).view(H, W).cpu().numpy(),
        "position": gb.position.tensor.view(torch.float32).view(H, W, 4).cpu().numpy(),
        "normal": gb.normal.tensor.view(torch.float16).view(H, W, 4).float().cpu().numpy(),
        "albedo": gb.albedo.tensor.view(H, W, 4).cpu().numpy(),
        "mro": gb.mro.tensor.view(H, W, 4).cpu().numpy(),
    }, gb


def _compare_gbuffer(got, want, exact=True, max_bad=0):
    cov = want["tri"] != NONE
    if exact:
        assert np.array_equal(got["depth"].view(np.uint32), want["depth"].view(np.uint32))
        same = np.ones_like(cov)
    else:
        same = got["depth"].view(np.uint32) == want["depth"].view(np.uint32)
        assert (~same).sum() <= max_bad, "%d pixels differ in depth" % (~same).sum()
    m = cov & same
    assert np.abs(got["position"][m] - want["position"][m]).max() <= 1e-4 * max(1.0, np.abs(want["position"][m]).max())
    assert np.abs(got["normal"][m] - want["normal"][m]).max() <= 2e-3  # RGBA16F storage
    assert np.abs(got["albedo"][m].astype(int) - want["albedo"][m].astype(int)).max() <= 1
    assert np.abs(got["mro"][m].astype(int) - want["mro"][m].astype(int)).max() <= 1
    e = ~cov & same
    assert (got["albedo"][e] == 0).all() and (got["normal"][e] == 0).all() and (got["position"][e] == 0).all()


def _textured_scene():
    sp = model.uv_sphere(1.0, (0.0, 0.0, 0.0), 24, 48)
    sp.material = model.MaterialData(baseColorFactor=(1.0, 0.9, 0.8, 1.0), metallicFactor=0.7, roughnessFactor=0.9, normalScale=0.8)
    sp.material.baseTexture = model.checker_texture(64, 8)
    rng = np.random.default_rng(5)
    nm = np.zeros((32, 32, 4), np.uint8)
    nm[..., :2] = rng.integers(96, 160, (32, 32, 2))
    nm[..., 2:] = 255
    sp.material.normalTexture = model.TextureData.from_rgba8(nm, model.sampler_word(model.WRAP_MIRROR, model.WRAP_CLAMP))
    mr = rng.integers(0, 256, (16, 16, 4)).astype(np.uint8)
    sp.material.metallicRoughnessTexture = model.TextureData.from_rgba8(mr, model.sampler_word(mip_mode=model.MIP_NEAREST))
    floor = model.quad([(-4, -1, 4), (4, -1, 4), (4, -1, -4), (-4, -1, -4)], 6.0)
    floor.material.baseTexture = model.checker_texture(128, 16, sampler=model.sampler_word(srgb=True))
    floor.model = np.array([[1, 0, 0, 0.2], [0, 1, 0, 0.0], [0, 0, 1, -0.3], [0, 0, 0, 1]], np.float32)
    return [sp, floor]


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(160, 90), (257, 131)])
def test_gpu_gbuffer_matches_the_restatement(ctx_fast, oracle, size):
    W, H = size
    g, proj, view = _camera(W, H, pos=(0.3, 0.8, 3.5), yaw=0.1, pitch=-0.2)
    prims = _textured_scene()
    want = oracle.draw_gbuffer(proj, view, prims, W, H)
    got, _ = _gpu_gbuffer(ctx_fast, g, prims, W, H)
    assert 0.3 < (want["tri"] != NONE).mean() < 1.0
    _compare_gbuffer(got, want)


@pytest.mark.gpu
def test_gpu_triangle_soup_and_eye_plane_crossings(ctx_fast, oracle):
    """Random triangles all around (and through) the camera, degenerate ones included: depth must match bit for bit."""
    W, H = 192, 108
    rng = np.random.default_rng(17)
    pts = rng.uniform(-6, 6, (600, 3, 3)).astype(np.float32)
    pts[::50, 2] = pts[::50, 1]  # degenerate: two equal vertices
    soup = _tri_prim(pts.reshape(-1, 3))
    floor = model.quad([(-50, -1.5, 50), (50, -1.5, 50), (50, -1.5, -50), (-50, -1.5, -50)])
    g, proj, view = _camera(W, H, pos=(0.0, 0.0, 0.0))
    prims = [soup, floor, _tri_prim(pts[::-1].reshape(-1, 3), front_cw=True)]
    want = oracle.draw_gbuffer(proj, view, prims, W, H)
    got, _ = _gpu_gbuffer(ctx_fast, g, prims, W, H)
    assert (want["tri"] != NONE).mean() > 0.5
    _compare_gbuffer(got, want)


@pytest.mark.gpu
def test_gpu_alpha_cutout(ctx_fast, oracle):
    W, H = 128, 96
    q = model.quad([(-1.5, -1, 0), (1.5, -1, 0), (1.5, 1, 0), (-1.5, 1, 0)], 1.0)
    q.material.baseTexture = model.checker_texture(32, 8, alpha_holes=True, sampler=model.sampler_word(srgb=True, mag_nearest=True, mip_mode=model.MIP_NONE))
    back = model.quad([(-3, -2, -1), (3, -2, -1), (3, 2, -1), (-3, 2, -1)])
    g, proj, view = _camera(W, H, pos=(0.2, 0.1, 2.5))
    want = oracle.draw_gbuffer(proj, view, [q, back], W, H)
    got, _ = _gpu_gbuffer(ctx_fast, g, [q, back], W, H)
    ids = want["tri"]
    assert (ids < 2).mean() > 0.1 and ((ids >= 2) & (ids != NONE)).mean() > 0.3  # holes show the quad behind
    # the alpha test compares a filtered value with the cutoff; contraction differences may move a handful of pixels
    _compare_gbuffer(got, want, exact=False, max_bad=int(0.001 * W * H))


@pytest.mark.gpu
def test_gpu_shadow_cubes_match_the_restatement(ctx_fast, oracle):
    import torch

    from althea_b200 import engine
    res = 64
    prims = _shadow_scene()
    lights = engine.PointLightCollection(ctx_fast, 3, shadow_res=res)
    pos = [(0.3, 2.0, 0.7), (-2.0, 0.5, 1.0), (4.0, 3.0, -3.0)]
    for i, p in enumerate(pos):
        lights.setLight(i, engine.PointLight(p, (10.0, 10.0, 10.0)))
    up = model.UploadedModel(ctx_fast, prims)
    lights.drawShadowMaps([up])
    torch.cuda.synchronize()
    got = lights.shadow_map.tensor.view(torch.float32).view(3, 6, res, res).cpu().numpy()
    pc = model.point_light_constants()
    views = np.array([list(pc.views[f]) for f in range(6)], np.float32)
    lt = np.zeros((3, 8), np.float32)
    lt[:, :3] = pos
    want = oracle.draw_shadow_cubes(lt, list(pc.projection), views, prims, res)
    assert (want < 1).mean() > 0.9
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_produced_gbuffer_feeds_the_deferred_chain(ctx_fast):
    """Producer -> consumer at 4K: G-buffer and shadow cubes rasterised from meshes, then SSR capture, glossy mips, SSAO and
    deferred shading on them; determinism, finiteness, coverage consistent between attachments."""
    import torch

    from althea_b200 import _capi, engine
    from helpers import FrameData, GpuFrame
    W, H = 3840, 2160
    prims = _textured_scene() + [model.uv_sphere(0.5, (1.6, -0.5, 0.8), 32, 64)]
    g = scene.make_uniforms(W, H, pos=(0.3, 0.8, 3.5), yaw=0.1, pitch=-0.2, light_count=2)
    up = model.UploadedModel(ctx_fast, prims)
    gb = engine.GBufferResources(ctx_fast, W, H)
    gpass = engine.SceneToGBufferPass(ctx_fast)
    gpass.draw(g, up, gb)
    torch.cuda.synchronize()
    d1 = gb.depth.tensor.clone()
    n1 = gb.normal.tensor.clone()
    gpass.draw(g, up, gb)
    torch.cuda.synchronize()
    assert torch.equal(d1, gb.depth.tensor) and torch.equal(n1, gb.normal.tensor)  # atomicMin resolution is order-independent
    depth = gb.depth.tensor.view(torch.float32).view(H, W)
    pos = gb.position.tensor.view(torch.float32).view(H, W, 4)
    nrm = gb.normal.tensor.view(torch.float16).view(H, W, 4).float()
    cov = depth < 1.0
    assert 0.3 < float(cov.float().mean()) < 1.0
    assert bool(((pos[..., 3] == 1.0) == cov).all()) and bool(((nrm[..., 3] > 0) == cov).all())
    assert float((nrm[cov][:, :3].norm(dim=1) - 1).abs().max()) < 2e-3
    lights = engine.PointLightCollection(ctx_fast, 2, shadow_res=256)
    lights.setLight(0, engine.PointLight((2.0, 3.0, 2.0), (30.0, 28.0, 25.0)))
    lights.setLight(1, engine.PointLight((-2.5, 1.5, 1.0), (10.0, 14.0, 20.0)))
    lights.drawShadowMaps([up])
    small = GpuFrame(ctx_fast, FrameData("scene", 32, 18, n_lights=0))  # its IBL set
    ssr = engine.ScreenSpaceReflection(ctx_fast, W, H)
    dp = engine.DeferredPass(ctx_fast, W, H, _capi.FORMAT_R16G16B16A16_SFLOAT)
    ssr.captureReflection(g, gb, small.ibl, lights)
    ssr.convolveReflectionBuffer()
    dp.draw(g, gb, small.ibl, lights, ssr, 0)
    torch.cuda.synchronize()
    col = dp.colorTarget.tensor.view(torch.float16).view(H, W, 4).float()
    assert bool(torch.isfinite(col).all())
    assert float(col[..., :3][cov].mean()) > 0.02
    shadow = lights.shadow_map.tensor.view(torch.float32)
    assert 0.05 < float((shadow < 1).float().mean()) <= 1.0  # an open scene: most directions see nothing


# ---- BASELINE configs[0]: the DamagedHelmet frame at 1280 x 720 ------------------------------------------------------------
def test_helmet_fixture_answer_is_reproducible(oracle):
    """The committed answer (CRC of the restatement's depth image and triangle ids at 1280 x 720) is what the restatement
    computes from the committed mesh and camera block on this machine too."""
    import zlib

    from helmet_fixture import CAMERA, GOLDEN, SIZE, helmet_primitives
    d = np.load(GOLDEN)
    prims = helmet_primitives()
    assert prims[0].triangle_count == 46356 // 3
    W, H = SIZE
    g = CAMERA()
    want = oracle.draw_gbuffer(list(g.projection), list(g.view), prims, W, H)
    assert zlib.crc32(want["depth"].tobytes()) == int(d["depth_crc"][0])
    assert zlib.crc32(want["tri"].tobytes()) == int(d["tri_crc"][0])
    assert int((want["tri"] != NONE).sum()) == int(d["coverage"][0])
    assert np.abs(want["albedo"][4::8, 4::8].astype(int) - d["albedo_thumb"].astype(int)).max() <= 1


@pytest.mark.gpu
def test_gpu_helmet_frame(ctx_fast, ctx_parity, oracle, tmp_path):
    """BASELINE configs[0] end to end on the GPU at its own shape: G-buffer of the helmet at 1280 x 720 (depth CRC equal to the
    committed answer, attributes against the live restatement), omni shadow cubes of two lights rendered from the same mesh,
    then SSR, glossy mips, SSAO and deferred shading in the PARITY build with a real IBL set, against the frame oracle fed with
    the same G-buffer and cubes: hit mask and AO counts bit for bit, colour to the bar (max abs error reported)."""
    import zlib

    import torch

    from althea_b200 import _capi, engine
    from helmet_fixture import CAMERA, GOLDEN, SIZE, helmet_primitives
    from helpers import FrameData, GpuFrame, half_to_float
    d = np.load(GOLDEN)
    prims = helmet_primitives()
    W, H = SIZE
    g = CAMERA()
    got, gb_fast = _gpu_gbuffer(ctx_fast, g, prims, W, H)
    assert zlib.crc32(np.ascontiguousarray(got["depth"]).tobytes()) == int(d["depth_crc"][0])
    assert int((got["depth"] < 1).sum()) == int(d["coverage"][0])
    want = oracle.draw_gbuffer(list(g.projection), list(g.view), prims, W, H)
    _compare_gbuffer(got, want)
    cov = want["tri"] != NONE
    # ---- the deferred chain on that G-buffer, parity build
    ctx = ctx_parity
    up = model.UploadedModel(ctx, prims)
    gb = engine.GBufferResources(ctx, W, H)
    engine.SceneToGBufferPass(ctx).draw(g, up, gb)
    n_lights, res = 2, 128
    g.lightCount = n_lights
    lights = engine.PointLightCollection(ctx, n_lights, res, True)
    lights.setLight(0, engine.PointLight((1.5, 1.2, 2.0), (30.0, 26.0, 20.0)))
    lights.setLight(1, engine.PointLight((-2.0, 0.5, 1.0), (8.0, 12.0, 20.0)))
    lights.updateResource()
    lights.drawShadowMaps([up])
    torch.cuda.synchronize()
    tiny = FrameData("scene", 32, 18, n_lights=0)  # for its IBL set: the golden environment, its mips, the reference's LUT
    ibl = GpuFrame(ctx, tiny).ibl
    ssr = engine.ScreenSpaceReflection(ctx, W, H)
    dp = engine.DeferredPass(ctx, W, H, _capi.FORMAT_R32G32B32A32_SFLOAT)
    ssr.captureReflection(g, gb, ibl, lights)
    ssr.convolveReflectionBuffer()
    dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP)
    torch.cuda.synchronize()
    arr = lambda img, dt, c: img.tensor.view(dt).view(H, W, c).cpu().numpy()  # noqa: E731
    position, depth = arr(gb.position, torch.float32, 4), gb.depth.tensor.view(torch.float32).view(H, W).cpu().numpy()
    normal = gb.normal.tensor.view(torch.int16).view(H, W, 4).cpu().numpy().view(np.uint16)
    albedo, mro = arr(gb.albedo, torch.uint8, 4), arr(gb.mro, torch.uint8, 4)
    cubes = lights.shadow_map.tensor.view(torch.float32).view(n_lights, 6, res, res).cpu().numpy()
    assert 0.01 < float((cubes < 1).mean()) < 0.9                       # the helmet shadows part of every cube
    og = oracle.GlobalUniforms.from_buffer_copy(bytes(g))
    fr = oracle.Frame(og, W, H, position, depth, normal, albedo, mro, tiny.env, tiny.pre, tiny.pre_size, 5, tiny.irr, tiny.lut, lights._lights.copy(), cubes, res)
    refl, hit, _ = oracle.ssr_capture(fr)
    chain = oracle.glossy_convolve(refl)
    ao = oracle.ssao(fr)
    col = oracle.deferred_shade(fr, chain, 5, oracle.SKIP_TONEMAP, ao)
    got_refl = ssr.getReflectionBuffer().image.level_numpy(0).view(np.uint16).reshape(H, W, 4)
    got_hit = half_to_float(got_refl)[..., 3] != 0
    assert hit[cov].mean() > 0.01, "the helmet must reflect itself somewhere"
    assert np.array_equal(got_hit, hit != 0)
    got_ao = dp.aoCounts.tensor.view(H, W).cpu().numpy()
    assert np.array_equal(got_ao, ao)
    assert (ao[cov] <= 24).all() and (ao[~cov] == 255).all() and (ao[cov] > 0).mean() > 0.2
    got_col = dp.colorTarget.tensor.view(torch.float32).view(H, W, 4).cpu().numpy()
    err = np.abs(got_col - col)
    ok = (err <= 1e-3 * np.maximum(1.0, np.abs(col))).all(axis=-1)
    assert ok.mean() >= 0.9995, "colour off the bar on %.4f %% of pixels, max abs error %.3g" % (100 * (1 - ok.mean()), err.max())
    assert float(np.percentile(err, 99.9)) <= 1e-3, "99.9th percentile of the absolute colour error on the RGBA32F target: %.3g" % np.percentile(err, 99.9)
    # the golden dump of configs[0]: the frame as the EXR file Utilities::saveExr writes (Src/Utilities.cpp:258-271), read back by
    # the reference's own tinyexr where oracle/_ref carries it
    import ctypes as C

    from althea_b200 import hdr_cache
    exr = str(tmp_path / "helmet_1280x720.exr")
    hdr_cache.save_exr(exr, got_col)
    assert os.path.getsize(exr) > W * H * 16
    ref_lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libtinyexr_ref.so")
    if os.path.exists(ref_lib):
        lib = C.CDLL(ref_lib)
        lib.ref_load_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        back = np.zeros_like(got_col)
        ww, hh = C.c_int(), C.c_int()
        assert lib.ref_load_exr(os.fsencode(exr), back.ctypes.data, back.size, C.byref(ww), C.byref(hh)) == 0
        assert (ww.value, hh.value) == (W, H) and np.array_equal(back.view(np.uint32), np.ascontiguousarray(got_col).view(np.uint32))


@pytest.mark.gpu
def test_gpu_work_list_overflow_is_retried(ctx_fast):
    """More tile work items than the initial list holds (hundreds of stacked full-screen quads at 4K): the shim notices the
    sticky overflow flag after the pass and redoes the call with a larger list; the result is the nearest quad everywhere."""
    W, H = 3840, 2160
    g = scene.make_uniforms(W, H, pos=(0.0, 0.0, 5.0), yaw=0.0, pitch=0.0)
    quads = [model.quad([(-40, -30, -k * 0.01), (40, -30, -k * 0.01), (40, 30, -k * 0.01), (-40, 30, -k * 0.01)]) for k in range(330)]
    assert 330 * 2 * ((W + 63) // 64) * ((H + 63) // 64) > (1 << 20) + 660
    got_all, _ = _gpu_gbuffer(ctx_fast, g, quads, W, H)
    got_first, _ = _gpu_gbuffer(ctx_fast, g, quads[:1], W, H)
    assert (got_first["depth"] < 1).all()
    assert np.array_equal(got_all["depth"].view(np.uint32), got_first["depth"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_record_and_work_lists_grow_on_demand(lib_built, oracle, monkeypatch):
    """The record and tile-work lists start small and grow when a pass overflows them (sticky demand counters, the call is
    redone). With the lists forced to 8 entries the G-buffer and the shadow cubes must still match the restatement exactly."""
    import torch

    from althea_b200 import engine
    monkeypatch.setenv("ALTHEA_RASTER_INITIAL_LISTS", "8")
    ctx = engine.Context(0)  # a fresh context: its scratch starts at the forced size
    try:
        W, H = 160, 90
        g, proj, view = _camera(W, H, pos=(0.3, 0.8, 3.5), yaw=0.1, pitch=-0.2)
        prims = _textured_scene()
        want = oracle.draw_gbuffer(proj, view, prims, W, H)
        got, _ = _gpu_gbuffer(ctx, g, prims, W, H)
        _compare_gbuffer(got, want)
        res = 64
        room = _shadow_scene()
        lights = engine.PointLightCollection(ctx, 2, shadow_res=res)
        pos = [(0.3, 2.0, 0.7), (-2.0, 0.5, 1.0)]
        for i, p in enumerate(pos):
            lights.setLight(i, engine.PointLight(p, (10.0, 10.0, 10.0)))
        lights.drawShadowMaps([model.UploadedModel(ctx, room)])
        torch.cuda.synchronize()
        cubes = lights.shadow_map.tensor.view(torch.float32).view(2, 6, res, res).cpu().numpy()
        pc = model.point_light_constants()
        lt = np.zeros((2, 8), np.float32)
        lt[:, :3] = pos
        want_c = oracle.draw_shadow_cubes(lt, list(pc.projection), np.array([list(pc.views[f]) for f in range(6)], np.float32), room, res)
        assert np.array_equal(cubes.view(np.uint32), want_c.view(np.uint32))
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_translucent_layers_blend_like_the_restatement(ctx_fast, oracle):
    """Alpha blending of the attachments on the GPU (depth winner + peeled layers) against the in-order restatement: translucent
    over opaque, translucent drawn first, three translucent layers over an opaque one (the four layers the peel keeps), and a
    translucent textured sphere in front of a textured floor."""
    W, H = 96, 64
    g, proj, view = _camera(W, H, pos=(0.1, 0.2, 2.5), yaw=0.05, pitch=-0.05)

    def layer(z, rgba, size=3.0, **kw):
        q = model.quad([(-size, -size, z), (size, -size, z), (size, size, z), (-size, size, z)])
        q.material = model.MaterialData(baseColorFactor=rgba, **kw)
        return q

    back = layer(-1.0, (1.0, 0.0, 0.0, 1.0), metallicFactor=1.0, roughnessFactor=1.0)
    front = layer(0.0, (0.0, 1.0, 0.0, 0.6), size=1.0, metallicFactor=0.0, roughnessFactor=0.5)
    mid = layer(-0.5, (0.0, 0.0, 1.0, 0.5), size=1.5)
    mid2 = layer(-0.25, (1.0, 1.0, 0.0, 0.7), size=1.2)
    sp = model.uv_sphere(0.7, (0.2, 0.1, 0.3), 16, 32)
    sp.material = model.MaterialData(baseColorFactor=(1.0, 1.0, 1.0, 0.75), metallicFactor=0.3, roughnessFactor=0.6)
    sp.material.baseTexture = model.checker_texture(64, 8)
    floor = _textured_scene()[1]
    for prims in ([back, front], [front, back], [back, mid, front], [back, mid, mid2, front], [mid, front], [floor, back, sp]):
        want = oracle.draw_gbuffer(proj, view, prims, W, H)
        got, _ = _gpu_gbuffer(ctx_fast, g, prims, W, H)
        _compare_gbuffer(got, want)
    mixed = oracle.draw_gbuffer(proj, view, [back, front], W, H)
    assert ((mixed["albedo"][..., 0] > 0) & (mixed["albedo"][..., 1] > 0)).any()  # red showing through green somewhere


def _write_gltf_json(path, attrs_arrays, extra=None, idx=None):
    """A .gltf with one embedded base64 buffer: attrs_arrays = {name: (array, componentType, type)}."""
    import base64
    import json
    blob, views, accessors, attrs = b"", [], [], {}
    for name, (arr, ctype, typ) in attrs_arrays.items():
        data = np.ascontiguousarray(arr).tobytes()
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)})
        blob += data + b"\0" * ((-len(data)) % 4)
        accessors.append({"bufferView": len(views) - 1, "componentType": ctype, "count": len(arr), "type": typ})
        attrs[name] = len(accessors) - 1
    prim = {"attributes": {k: v for k, v in attrs.items() if k != "_indices" and k != "_ibm"}}
    if "_indices" in attrs:
        prim["indices"] = attrs["_indices"]
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}], "meshes": [{"primitives": [prim]}],
         "bufferViews": views, "accessors": accessors,
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    if extra:
        extra(g, attrs)
    with open(path, "w") as f:
        json.dump(g, f)


def test_loader_edge_cases_follow_primitive_cpp(tmp_path):
    """Four corners where a tidy loader and Src/Primitive.cpp disagree; the loader follows the reference:
    TEXCOORD sets stop at the first missing index (:126-134); TANGENT without NORMAL leaves tangent and bitangent zero (:336-345);
    u8 / u16 WEIGHTS_0 are divided by 255 / 65535 whatever the accessor's `normalized` says (:354-359); a primitive is drawn with
    global * inverseBindPose of its own node (Model.cpp:349)."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    uv0 = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    uv2 = uv0 + 5.0
    tan = np.tile(np.array([[1, 0, 0, 1]], np.float32), (3, 1))
    path = str(tmp_path / "gap.gltf")
    _write_gltf_json(path, {"POSITION": (pos, 5126, "VEC3"), "TANGENT": (tan, 5126, "VEC4"), "TEXCOORD_0": (uv0, 5126, "VEC2"),
                            "TEXCOORD_2": (uv2, 5126, "VEC2")})
    (p,) = model.load_gltf(path)
    v = p.vertices
    assert np.array_equal(v[:, 12:14], uv0) and not v[:, 14:20].any()          # set 2 is behind a gap: not loaded, not moved to slot 1
    assert np.allclose(v[:, 9:12], [0, 0, 1]) and not v[:, 3:9].any()          # flat normal generated; tangent frame left at zero
    # skinned: one joint translated by (1, 0, 0), u8 weights (255, 0, 0, 0) WITHOUT the normalized flag
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (3, 1))
    joints = np.zeros((3, 4), np.uint8)
    weights = np.tile(np.array([[255, 0, 0, 0]], np.uint8), (3, 1))
    ibm = np.eye(4, dtype=np.float32).reshape(1, 16)

    def skin(g, attrs):
        g["nodes"] = [{"mesh": 0, "skin": 0}, {"translation": [1.0, 0.0, 0.0]}]
        g["scenes"][0]["nodes"] = [0, 1]
        g["skins"] = [{"joints": [1], "inverseBindMatrices": attrs["_ibm"]}]

    path = str(tmp_path / "skin.gltf")
    _write_gltf_json(path, {"POSITION": (pos, 5126, "VEC3"), "NORMAL": (nrm, 5126, "VEC3"), "TANGENT": (tan, 5126, "VEC4"), "TEXCOORD_0": (uv0, 5126, "VEC2"),
                            "JOINTS_0": (joints, 5121, "VEC4"), "WEIGHTS_0": (weights, 5121, "VEC4"), "_ibm": (ibm, 5126, "MAT4")}, extra=skin)
    (p,) = model.load_gltf(path)
    assert np.allclose(p.vertices[:, 0:3], pos + [1, 0, 0])                    # weight 1, not 255
    assert np.allclose(p.vertices[:, 20:24], [1, 0, 0, 0])
    # a mesh on a node that is itself a joint with a non-identity inverse bind pose: drawn with global * inverseBind
    ibm2 = np.eye(4, dtype=np.float32)
    ibm2[3, 0] = -3.0  # column-major storage: translation by (-3, 0, 0)

    def joint_mesh(g, attrs):
        g["nodes"] = [{"mesh": 0, "translation": [1.0, 2.0, 0.0]}, {"skin": 0}]
        g["scenes"][0]["nodes"] = [0, 1]
        g["skins"] = [{"joints": [0], "inverseBindMatrices": attrs["_ibm"]}]

    path = str(tmp_path / "jointmesh.gltf")
    _write_gltf_json(path, {"POSITION": (pos, 5126, "VEC3"), "NORMAL": (nrm, 5126, "VEC3"), "TANGENT": (tan, 5126, "VEC4"), "_ibm": (ibm2.reshape(1, 16), 5126, "MAT4")},
                     extra=joint_mesh)
    (p,) = model.load_gltf(path)
    assert np.allclose(p.model[:3, 3], [1.0 - 3.0, 2.0, 0.0])
