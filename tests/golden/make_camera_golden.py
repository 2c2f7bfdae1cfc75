"""Generates tests/golden/camera_ref.npz. Run in the BUILD container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_camera_golden.py

Outputs of the REFERENCE's own Camera class (Src/Camera.cpp compiled where it lies against its vendored GLM into
oracle/_ref/libcamera_ref.so): the PointLightConstants matrices of Src/PointLight.cpp:72-118 and a set of camera poses."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def poses():
    rs = np.random.default_rng(21)
    fixed = [(60.0, 16 / 9, (0.6, 0.35, 2.4), 0.25, -0.12), (60.0, 16 / 9, (0.0, 2.0, 6.0), 0.0, -0.25),
             (90.0, 1.0, (0.0, 0.0, 0.0), np.pi, np.pi / 2), (90.0, 1.0, (0.0, 0.0, 0.0), np.pi, -np.pi / 2),
             (45.0, 4 / 3, (1.0, 2.0, 3.0), 3.0, 3.2), (45.0, 4 / 3, (1.0, 2.0, 3.0), -3.0, -3.2)]
    rand = [(float(rs.uniform(20, 120)), float(rs.uniform(0.5, 2.5)), tuple(rs.uniform(-50, 50, 3)), float(rs.uniform(-7, 7)),
             float(rs.uniform(-3.5, 3.5))) for _ in range(58)]
    return fixed + rand


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcamera_ref.so"))
    plc = np.zeros((14, 16), np.float32)
    ref.ref_point_light_constants(plc.ctypes.data_as(C.c_void_p))
    F = C.c_float
    params, mats = [], []
    for fov, aspect, pos, yaw, pitch in poses():
        p = np.array(pos, np.float32)
        o = [np.zeros(16, np.float32) for _ in range(3)]
        ref.ref_camera(F(fov), F(aspect), F(0.01), F(1000.0), p.ctypes.data_as(C.c_void_p), F(yaw), F(pitch),
                       *[x.ctypes.data_as(C.c_void_p) for x in o])
        params.append([fov, aspect, *p, yaw, pitch])
        mats.append(np.stack(o))
    np.savez_compressed(os.path.join(HERE, "camera_ref.npz"), point_light_constants=plc, params=np.array(params, np.float32),
                        matrices=np.stack(mats))
    print(plc.shape, np.stack(mats).shape)


if __name__ == "__main__":
    sys.exit(main())
