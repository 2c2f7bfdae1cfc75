"""Synthetic inputs for the deferred path: camera blocks, S-rand / S-scene G-buffers, point lights with
ray-cast omni shadow cubes, a procedural HDR environment. Pure torch, device-agnostic (CPU for the parity tests, CUDA for
bench.py), deterministic: all randomness is a counter-based splitmix64 hash, so a (seed, view, pixel) triple gives the
same value on every device.

These are INPUT generators (test / bench plumbing); nothing here is on the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from ._capi import GlobalUniforms

SEED = 0xA17EA
NEAR, FAR = 0.01, 1000.0  # CameraController.cpp:13-14; hard-coded again in Misc/ReconstructPosition.glsl:6-7


# ---- counter-based RNG ------------------------------------------------------------------------------------------
def _i64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x: torch.Tensor, n: int) -> torch.Tensor:
    return (x >> n) & ((1 << (64 - n)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def hash_uniform(index: torch.Tensor, stream: int, seed: int = SEED, view: int = 0) -> torch.Tensor:
    """U[0,1) float32 per element of `index` (int64), independent across `stream` ids."""
    key = _i64(seed ^ (view << 40) ^ (stream << 52))
    z = splitmix64(splitmix64(index ^ key))
    return (_lsr(z, 40).to(torch.float32) * (1.0 / 16777216.0))


# ---- camera -----------------------------------------------------------------------------------------------------
def perspective(fov_deg: float, aspect: float, near: float = NEAR, far: float = FAR) -> np.ndarray:
    """glm::perspective RH, depth 0..1, then [1][1] *= -1 (Src/Camera.cpp:90-103). Returns row-major 4x4 float64."""
    t = math.tan(math.radians(fov_deg) / 2.0)
    m = np.zeros((4, 4), np.float64)
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = -1.0 / t
    m[2, 2] = far / (near - far)
    m[3, 2] = -1.0
    m[2, 3] = -(far * near) / (far - near)
    return m


def camera_transform(pos, yaw: float, pitch: float) -> np.ndarray:
    """Src/Camera.cpp:49-65: columns (xAxis, yAxis, zAxis (backward), position)."""
    cp = math.cos(pitch)
    z = np.array([math.sin(yaw) * cp, -math.sin(pitch), math.cos(yaw) * cp])
    x = np.cross(np.array([0.0, 1.0, 0.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, np.asarray(pos, np.float64)
    return m


def make_uniforms(width: int, height: int, pos=(0.0, 0.0, 0.0), yaw: float = 0.0, pitch: float = 0.0, fov_deg: float = 60.0,
                  light_count: int = 0, exposure: float = 0.6) -> GlobalUniforms:
    proj = perspective(fov_deg, width / height)
    xf = camera_transform(pos, yaw, pitch)
    view = np.linalg.inv(xf)
    g = GlobalUniforms()

    def put(field, m):
        arr = np.asarray(m, np.float64).T.astype(np.float32).reshape(-1)  # column-major
        for i in range(16):
            field[i] = float(arr[i])

    put(g.projection, proj)
    put(g.inverseProjection, np.linalg.inv(proj))
    put(g.view, view)
    put(g.prevView, view)
    put(g.inverseView, xf)
    put(g.prevInverseView, xf)
    g.lightCount = light_count
    g.exposure = exposure
    g.time = 0.0
    return g


def _mat(field) -> np.ndarray:
    return np.array(list(field), np.float64).reshape(4, 4).T  # back to row-major


# ---- G-buffers --------------------------------------------------------------------------------------------------
@dataclass
class GBufferData:
    """Raw attachment contents, torch tensors on one device, layouts as the reference allocates them."""
    width: int
    height: int
    position: torch.Tensor  # (H, W, 4) float32, .a == 0 => empty
    depth: torch.Tensor     # (H, W) float32
    normal: torch.Tensor    # (H, W, 4) float16
    albedo: torch.Tensor    # (H, W, 4) uint8
    mro: torch.Tensor       # (H, W, 4) uint8

    def numpy(self):
        return dict(position=self.position.cpu().numpy(), depth=self.depth.cpu().numpy(),
                    normal=self.normal.cpu().view(torch.int16).numpy().view(np.uint16), albedo=self.albedo.cpu().numpy(),
                    mro=self.mro.cpu().numpy())


def _pixel_rays(g: GlobalUniforms, W: int, H: int, device):
    """World-space unit ray through each pixel centre + camera position (float32 tensors)."""
    inv_proj = torch.tensor(_mat(g.inverseProjection), dtype=torch.float64, device=device)
    inv_view = torch.tensor(_mat(g.inverseView), dtype=torch.float64, device=device)
    xs = (torch.arange(W, device=device, dtype=torch.float64) + 0.5) / W * 2.0 - 1.0
    ys = (torch.arange(H, device=device, dtype=torch.float64) + 0.5) / H * 2.0 - 1.0
    ndc_y, ndc_x = torch.meshgrid(ys, xs, indexing="ij")
    clip = torch.stack([ndc_x, ndc_y, torch.zeros_like(ndc_x), torch.ones_like(ndc_x)], -1)
    eye = clip @ inv_proj.T
    d = eye[..., :3] @ inv_view[:3, :3].T
    d = d / d.norm(dim=-1, keepdim=True)
    return d, inv_view[:3, 3]


def _depth_raw(dist: torch.Tensor) -> torch.Tensor:
    return FAR * (dist - NEAR) / ((FAR - NEAR) * dist)


def _pack(W, H, hit, pos, nrm, alb, metal, rough, dist) -> GBufferData:
    f32 = torch.float32
    a = hit.to(f32)
    position = torch.cat([pos.to(f32) * a[..., None], a[..., None]], -1)
    depth = torch.where(hit, _depth_raw(dist).to(f32), torch.ones_like(a))
    normal = torch.cat([nrm.to(f32) * a[..., None], a[..., None]], -1).to(torch.float16)
    alb8 = torch.cat([(alb.to(f32) * 255.0 + 0.5).floor().clamp(0, 255), torch.full_like(a, 255.0)[..., None]], -1) * a[..., None]
    mro8 = torch.stack([(metal.to(f32) * 255.0 + 0.5).floor(), (rough.to(f32) * 255.0 + 0.5).floor(), torch.full_like(a, 255.0),
                        torch.full_like(a, 255.0)], -1).clamp(0, 255) * a[..., None]
    return GBufferData(W, H, position.contiguous(), depth.contiguous(), normal.contiguous(), alb8.to(torch.uint8).contiguous(),
                       mro8.to(torch.uint8).contiguous())


def s_rand(g: GlobalUniforms, W: int, H: int, device="cpu", seed: int = SEED, view: int = 0) -> GBufferData:
    """S-rand (BASELINE C4/C5): per-pixel random depth / normal / albedo / MRO, 5 % empty pixels."""
    d, cam = _pixel_rays(g, W, H, device)
    idx = torch.arange(W * H, device=device, dtype=torch.int64).reshape(H, W)
    u = lambda s: hash_uniform(idx, s, seed, view).to(torch.float64)  # noqa: E731
    inv_view = torch.tensor(_mat(g.inverseView), dtype=torch.float64, device=device)
    fwd = -inv_view[:3, 2]
    eye_dist = 1.0 + 49.0 * u(0)               # distance along the view axis
    t = eye_dist / (d @ fwd)                   # ray parameter
    pos = cam + d * t[..., None]
    # normal: uniform on the hemisphere facing the camera
    z = u(1)
    phi = 2.0 * math.pi * u(2)
    r = torch.sqrt((1.0 - z * z).clamp_min(0.0))
    local = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], -1)
    nz = -d
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device=device).expand_as(nz)
    tx = torch.linalg.cross(up, nz)
    tx = tx / tx.norm(dim=-1, keepdim=True).clamp_min(1e-9)
    ty = torch.linalg.cross(nz, tx)
    nrm = tx * local[..., 0:1] + ty * local[..., 1:2] + nz * local[..., 2:3]
    alb = torch.stack([u(3), u(4), u(5)], -1)
    metal = (u(6) < 0.5).to(torch.float64)
    rough = 0.05 + 0.95 * u(7)
    hit = u(8) >= 0.05
    return _pack(W, H, hit, pos, nrm, alb, metal, rough, eye_dist)


@dataclass
class AnalyticScene:
    """S-scene: ground plane y = 0 plus spheres resting on it."""
    centers: torch.Tensor  # (N, 3) float64
    radii: torch.Tensor    # (N,)
    albedo: torch.Tensor   # (N, 3)
    metal: torch.Tensor    # (N,)
    rough: torch.Tensor    # (N,)


def make_scene(n_spheres: int = 64, seed: int = SEED, device="cpu") -> AnalyticScene:
    i = torch.arange(n_spheres, device=device, dtype=torch.int64)
    u = lambda s: hash_uniform(i, 100 + s, seed).to(torch.float64)  # noqa: E731
    radii = 0.3 + 0.9 * u(0)
    centers = torch.stack([-8.0 + 16.0 * u(1), radii, -14.0 + 16.0 * u(2)], -1)
    albedo = torch.stack([u(3), u(4), u(5)], -1)
    return AnalyticScene(centers, radii, albedo, (u(6) < 0.5).to(torch.float64), 0.05 + 0.95 * u(7))


def make_ring_scene(n_spheres: int = 160, seed: int = SEED, device="cpu", r_min: float = 3.5, r_max: float = 16.0) -> AnalyticScene:
    """Spheres on an annulus around the origin, so a camera at (0, 2, 0) sees similar content at every yaw (bench views)."""
    i = torch.arange(n_spheres, device=device, dtype=torch.int64)
    u = lambda s: hash_uniform(i, 300 + s, seed).to(torch.float64)  # noqa: E731
    radii = 0.3 + 0.9 * u(0)
    ang = 2.0 * math.pi * u(1)
    rad = torch.sqrt(r_min ** 2 + (r_max ** 2 - r_min ** 2) * u(2))
    centers = torch.stack([rad * torch.sin(ang), radii, rad * torch.cos(ang)], -1)
    albedo = torch.stack([u(3), u(4), u(5)], -1)
    return AnalyticScene(centers, radii, albedo, (u(6) < 0.5).to(torch.float64), 0.05 + 0.95 * u(7))


def _raycast(scene: AnalyticScene, origin: torch.Tensor, d: torch.Tensor):
    """Nearest hit of rays origin + t d (d unit). origin broadcastable to d. Returns t (inf = miss) and object id (-1 plane)."""
    inf = torch.full(d.shape[:-1], float("inf"), dtype=d.dtype, device=d.device)
    oy = origin[..., 1].expand(d.shape[:-1])
    t_plane = torch.where(d[..., 1] < -1e-9, -oy / d[..., 1].clamp_max(-1e-9), inf)
    t_plane = torch.where(t_plane > 1e-6, t_plane, inf)
    best, obj = t_plane, torch.full(d.shape[:-1], -1, dtype=torch.int64, device=d.device)
    for s in range(scene.centers.shape[0]):
        oc = origin - scene.centers[s]
        b = (oc * d).sum(-1)
        c = (oc * oc).sum(-1) - scene.radii[s] ** 2
        disc = b * b - c
        sq = torch.sqrt(disc.clamp_min(0.0))
        t0 = -b - sq
        t = torch.where((disc > 0) & (t0 > 1e-6), t0, inf)
        closer = t < best
        best = torch.where(closer, t, best)
        obj = torch.where(closer, torch.full_like(obj, s), obj)
    return best, obj


def s_scene(g: GlobalUniforms, W: int, H: int, scene: AnalyticScene, device="cpu") -> GBufferData:
    d, cam = _pixel_rays(g, W, H, device)
    t, obj = _raycast(scene, cam, d)
    hit = torch.isfinite(t) & (t < 200.0)
    t = torch.where(hit, t, torch.ones_like(t))
    pos = cam + d * t[..., None]
    is_sphere = obj >= 0
    oi = obj.clamp_min(0)
    nrm_s = (pos - scene.centers[oi]) / scene.radii[oi][..., None]
    nrm = torch.where(is_sphere[..., None], nrm_s, torch.tensor([0.0, 1.0, 0.0], dtype=pos.dtype, device=device).expand_as(pos))
    checker = ((torch.floor(pos[..., 0]) + torch.floor(pos[..., 2])) % 2 == 0).to(pos.dtype)
    ground_alb = (0.35 + 0.4 * checker)[..., None].expand_as(pos)
    alb = torch.where(is_sphere[..., None], scene.albedo[oi], ground_alb)
    metal = torch.where(is_sphere, scene.metal[oi], torch.zeros_like(t))
    rough = torch.where(is_sphere, scene.rough[oi], torch.full_like(t, 0.15))
    inv_view = torch.tensor(_mat(g.inverseView), dtype=torch.float64, device=device)
    eye_dist = ((pos - cam) @ (-inv_view[:3, 2]))
    return _pack(W, H, hit, pos, nrm, alb, metal, rough, eye_dist.clamp_min(NEAR * 2))


# ---- lights -----------------------------------------------------------------------------------------------------
def make_lights(n: int, seed: int = SEED, device="cpu", ring: bool = False) -> torch.Tensor:
    """(n, 8) float32 PointLight rows: positions U[-10,10]x[1,6]x[-10,10] (z shifted to sit over the test
    scene, or centred on the origin with ring=True), emission U[5,50]^3."""
    i = torch.arange(n, device=device, dtype=torch.int64)
    u = lambda s: hash_uniform(i, 200 + s, seed)  # noqa: E731
    out = torch.zeros(n, 8, dtype=torch.float32, device=device)
    out[:, 0] = -10.0 + 20.0 * u(0)
    out[:, 1] = 1.0 + 5.0 * u(1)
    out[:, 2] = (-10.0 + 20.0 * u(2)) if ring else (-14.0 + 18.0 * u(2))
    out[:, 4] = 5.0 + 45.0 * u(3)
    out[:, 5] = 5.0 + 45.0 * u(4)
    out[:, 6] = 5.0 + 45.0 * u(5)
    return out


def shadow_cubes(scene: AnalyticScene, lights: torch.Tensor, res: int = 256) -> torch.Tensor:
    """(n, 6, res, res) float32 = distance to the nearest surface / 1000 along the direction the CONSUMER associates with
    each texel (PBRMaterial.glsl:130 looks up q = (L.x,-L.y,-L.z), L = surface->light; so texel direction q maps to the
    world direction light->surface (-q.x, q.y, q.z)). 1.0 where nothing is hit."""
    device = lights.device
    n = lights.shape[0]
    c = (torch.arange(res, device=device, dtype=torch.float64) + 0.5) / res * 2.0 - 1.0
    tc, sc = torch.meshgrid(c, c, indexing="ij")
    one = torch.ones_like(sc)
    faces = [(one, -tc, -sc), (-one, -tc, sc), (sc, one, tc), (sc, -one, -tc), (sc, -tc, one), (-sc, -tc, -one)]
    out = torch.ones(n, 6, res, res, dtype=torch.float32, device=device)
    for li in range(n):
        origin = lights[li, :3].to(torch.float64)
        for f, (qx, qy, qz) in enumerate(faces):
            d = torch.stack([-qx, qy, qz], -1)
            d = d / d.norm(dim=-1, keepdim=True)
            t, _ = _raycast(scene, origin, d)
            out[li, f] = torch.where(torch.isfinite(t), (t / 1000.0), torch.ones_like(t)).clamp_max(1.0).to(torch.float32)
    return out


# ---- environment ------------------------------------------------------------------------------------------------
def procedural_env(W: int, H: int, device="cpu") -> torch.Tensor:
    """(H, W, 4) float32 equirect HDR sky: gradient + warm sun + a few bright panels (stands in for HDRI_Skybox/*.hdr)."""
    v = (torch.arange(H, device=device, dtype=torch.float32) + 0.5) / H
    u = (torch.arange(W, device=device, dtype=torch.float32) + 0.5) / W
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    pitch = (vv - 0.5) * math.pi
    yaw = (uu * 2.0 - 1.0) * math.pi
    d = torch.stack([torch.cos(pitch) * torch.cos(yaw), torch.sin(pitch), torch.cos(pitch) * torch.sin(yaw)], -1)
    sky = torch.stack([0.25 + 0.35 * vv, 0.35 + 0.35 * vv, 0.55 + 0.4 * vv], -1)
    ground = torch.tensor([0.22, 0.2, 0.18], device=device)
    base = torch.where((d[..., 1:2] < 0.0), sky, ground.expand_as(sky))
    sun_dir = torch.tensor([0.4, -0.6, -0.69], device=device)
    sun_dir = sun_dir / sun_dir.norm()
    cs = (d @ sun_dir).clamp(0.0, 1.0)
    sun = (cs ** 800.0)[..., None] * torch.tensor([120.0, 100.0, 70.0], device=device) + (cs ** 12.0)[..., None] * 0.8
    panels = torch.zeros_like(base)
    for k, (cy, cp, amp) in enumerate([(1.9, 0.3, 6.0), (-2.4, -0.2, 4.0), (0.3, 0.9, 3.0)]):
        m = ((yaw - cy).abs() < 0.25) & ((pitch - cp).abs() < 0.12)
        panels = panels + m[..., None].to(torch.float32) * amp
    rgb = base + sun + panels
    return torch.cat([rgb, torch.ones_like(rgb[..., :1])], -1).contiguous()
