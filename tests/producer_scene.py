"""The two small scenes of the producer pins (tests/test_shader_ref.py, tests/golden/make_shader_golden.py): a textured sphere over
a floor with a normal map, an sRGB base texture, mirrored / clamped / nearest-mip samplers, an alpha-cutout quad and a translucent
one (G-buffer pass), and a closed room with a sphere and the cutout quad (omni shadow pass)."""
import numpy as np

from althea_b200 import model, scene


def _camera(W, H, pos, yaw, pitch):
    g = scene.make_uniforms(W, H, pos=pos, yaw=yaw, pitch=pitch)
    return g, list(g.projection), list(g.view)


def _cutout():
    q = model.quad([(-1.5, -0.5, 1.2), (0.5, -0.5, 1.2), (0.5, 1, 1.2), (-1.5, 1, 1.2)], 1.0)
    q.material.baseTexture = model.checker_texture(32, 8, alpha_holes=True, sampler=model.sampler_word(srgb=True, mag_nearest=True, mip_mode=model.MIP_NONE))
    return q


def gbuffer_case():
    W, H = 128, 72
    g, proj, view = _camera(W, H, (0.3, 0.8, 3.5), 0.1, -0.2)
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
    glass = model.quad([(0.2, -0.8, 1.6), (1.6, -0.8, 1.6), (1.6, 0.9, 1.6), (0.2, 0.9, 1.6)])
    glass.material = model.MaterialData(baseColorFactor=(0.2, 0.9, 0.4, 0.55), metallicFactor=0.3, roughnessFactor=0.6)
    return proj, view, [sp, floor, _cutout(), glass], W, H


def shadow_case():
    pc = model.point_light_constants()
    views = np.array([list(pc.views[f]) for f in range(6)], np.float32)
    lights = np.zeros((2, 8), np.float32)
    lights[0, :3], lights[1, :3] = (0.3, 2.0, 0.7), (-2.0, 3.0, -1.0)
    sp = model.uv_sphere(0.8, (0.5, 0.6, -0.4), 10, 20)
    room = [model.quad(c) for c in ([(-6, -1, 6), (6, -1, 6), (6, -1, -6), (-6, -1, -6)], [(-6, 5, -6), (6, 5, -6), (6, 5, 6), (-6, 5, 6)],
                                    [(-6, -1, -5), (6, -1, -5), (6, 6, -5), (-6, 6, -5)], [(6, -1, 6), (-6, -1, 6), (-6, 6, 6), (6, 6, 6)],
                                    [(-6, -1, 6), (-6, -1, -6), (-6, 6, -6), (-6, 6, 6)], [(6, -1, -6), (6, -1, 6), (6, 6, 6), (6, 6, -6)])]
    return lights, list(pc.projection), views, [sp] + room + [_cutout()], 32
