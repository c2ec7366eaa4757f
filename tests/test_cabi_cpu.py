"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/svo_cuda.h declares,
mirrors the reference's struct layouts, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

from stereo_svo_slam_b200 import capi, synth


def test_library_exports_every_declared_symbol():
    syms = capi.declared_symbols()
    assert len(syms) >= 35
    lib = capi.lib()
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_struct_layouts_match_reference():
    # CameraSettings: 10 floats + 9 ints (src/include/stereo_slam_types.hpp:16-36); Pose: 6 floats (pose_manager.hpp:21-28)
    assert C.sizeof(capi.CameraSettings) == 19 * 4
    assert C.sizeof(capi.Pose) == 24
    names = [f[0] for f in capi.CameraSettings._fields_]
    assert names[:5] == ["baseline", "fx", "fy", "cx", "cy"] and names[10:12] == ["grid_height", "grid_width"]
    assert capi.KPINFO_DTYPE.itemsize == C.sizeof(capi.KeyPointInfo) == 56


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of every struct the Python binding mirrors, taken from include/svo_cuda.h by the C compiler"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    probes = [("svo_camera_settings", capi.CameraSettings, ["baseline", "grid_height", "min_pyramid_level_pose_estimation"]),
              ("svo_pose", capi.Pose, ["x", "rz"]),
              ("svo_keypoint_info", capi.KeyPointInfo, ["keyframe_id", "keypoint_index", "outlier_count", "kf_variance"]),
              ("svo_track_io", capi.TrackIO, ["n", "prev_kps2d", "pose_prior", "kps2d", "pose_aligned", "align_evals", "refine_evals",
                                              "klt_pts", "klt_iters", "keypoint_index"])]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "svo_cuda.h"', 'int main(void) {']
    for cname, _, fields in probes:
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({cname}, {f}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line, (cname, ctype, fields) in zip(out, probes):
        vals = line.split()
        assert vals[0] == cname and int(vals[1]) == C.sizeof(ctype), (cname, vals[1], C.sizeof(ctype))
        for v, f in zip(vals[2:], fields):
            assert int(v) == getattr(ctype, f).offset, (cname, f, v, getattr(ctype, f).offset)


def test_product_path_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "stereo_svo_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "ocv_prims" not in src and "svo_oracle" not in src, f


@pytest.mark.skipif(capi.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_device():
    cs = capi.CameraSettings(**synth.settings_dict("C3"))
    with pytest.raises(capi.SvoError) as e:
        capi.Context(cs, 752, 480)
    assert e.value.code == capi.SVO_ERR_NO_DEVICE
    from stereo_svo_slam_b200 import StereoSlam
    with pytest.raises(capi.SvoError):
        StereoSlam(cs, 752, 480)


def test_argument_validation_precedes_device_probe():
    cs = capi.CameraSettings(**synth.settings_dict("C3"))
    out = C.c_void_p()
    assert capi.lib().svo_ctx_create(None, 0, 752, 480, 0, C.byref(out)) == capi.SVO_ERR_INVALID
    bad = capi.CameraSettings(**dict(synth.settings_dict("C3"), window_size_opt_flow=33))
    assert capi.lib().svo_ctx_create(C.byref(bad), 0, 752, 480, 0, C.byref(out)) == capi.SVO_ERR_INVALID
    assert capi.lib().svo_ctx_destroy(None) == capi.SVO_ERR_INVALID
