"""CPU tests of the boundary: the shared library builds, loads, exports exactly the symbols include/althea_cuda.h
declares, the parameter blocks have the reference's byte layout, and nothing silently falls back when no GPU exists."""
import ctypes as C
import os
import re

import pytest

from althea_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "althea_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(althea_cuda_[a-z_]+)\s*\(", text))


def test_header_and_binding_agree():
    assert _declared_symbols() == set(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib_built):
    lib = C.CDLL(lib_built)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _capi.load().althea_cuda_abi_version() == 1


def test_parameter_block_layouts():
    # Include/Althea/GlobalUniforms.h:15-31 -> offsets per SURVEY.md App. B
    G = _capi.GlobalUniforms
    assert C.sizeof(G) == 416
    off = {f[0]: getattr(G, f[0]).offset for f in G._fields_}
    assert (off["projection"], off["inverseProjection"], off["view"], off["prevView"], off["inverseView"], off["prevInverseView"]) == (
        0, 64, 128, 192, 256, 320)
    assert (off["mouseUV"], off["lightCount"], off["lightBufferHandle"], off["time"], off["exposure"], off["inputMask"],
            off["frameCount"]) == (384, 392, 396, 400, 404, 408, 412)
    assert C.sizeof(_capi.GBuffer) == 40 and C.sizeof(_capi.IBL) == 32 and C.sizeof(_capi.Sync) == 40


def test_image_bytes(lib_built):
    lib = _capi.load()
    f = lib.althea_cuda_image_bytes
    assert f(_capi.FORMAT_R16G16B16A16_SFLOAT, 3840, 2160, 5, 1) == 8 * (3840 * 2160 + 1920 * 1080 + 960 * 540 + 480 * 270 + 240 * 135)
    assert f(_capi.FORMAT_R32_SFLOAT, 256, 256, 1, 96) == 4 * 256 * 256 * 96
    assert f(_capi.FORMAT_R32G32B32A32_SFLOAT, 4096, 2048, 13, 1) == 16 * sum(max(1, 4096 >> k) * max(1, 2048 >> k) for k in range(13))
    assert f(12345, 4, 4, 1, 1) == 0


def test_no_cpu_fallback_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from althea_b200 import engine
    with pytest.raises(engine.AltheaError, match="no CPU fallback|no CUDA"):
        engine.Context(0)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under althea_b200/ may reference it
    pkg = os.path.join(ROOT, "althea_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\"[^\"]*oracle", src, flags=re.M), os.path.join(dirpath, fn)


def test_header_is_plain_c_and_structs_match_the_binding(tmp_path):
    """include/althea_cuda.h is the boundary a C host binds: it must compile as C99 on its own, and the structs the Python
    binding mirrors must have the sizes and offsets the C compiler gives them."""
    import subprocess
    src = tmp_path / "probe.c"
    src.write_text('''#include <stddef.h>
#include <stdio.h>
#include "althea_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(althea_vertex), sizeof(althea_texture_ref), sizeof(althea_material), sizeof(althea_primitive),
         sizeof(althea_point_light_constants), offsetof(althea_primitive, model), offsetof(althea_primitive, material),
         offsetof(althea_material, baseTexture));
  printf("%zu %zu %zu %zu\\n", sizeof(althea_global_uniforms), sizeof(althea_point_light), sizeof(althea_gbuffer), sizeof(althea_sync));
  return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    P, M = _capi.Primitive, _capi.Material
    assert [int(v) for v in a.split()] == [_capi.VERTEX_BYTES, C.sizeof(_capi.TextureRef), C.sizeof(M), C.sizeof(P), C.sizeof(_capi.PointLightConstants),
                                           P.model.offset, P.material.offset, M.baseTexture.offset]
    assert [int(v) for v in b.split()] == [416, 32, C.sizeof(_capi.GBuffer), C.sizeof(_capi.Sync)]


def test_every_entry_point_is_documented():
    """INTEGRATION.md section 7 indexes every symbol of both headers (prefix-abbreviated rows such as `_destroy` count)."""
    import re
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for header, prefix in (("althea_cuda.h", "althea_cuda"), ("althea_host.h", "althea_host")):
        text = open(os.path.join(ROOT, "include", header)).read()
        names = set(re.findall(r"\b(%s_\w+)\s*\(" % prefix, text))
        for name in names:
            short = name[len(prefix):]
            assert name in doc or ("`%s`" % short) in doc, name
