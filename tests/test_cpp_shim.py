"""The C++ side of the drop-in boundary: include/glim_b200.h is valid C, the gtsam_points-compatible shims compile in both
modes (stand-alone, and -DGLIM_B200_WITH_GTSAM against signature stubs -- GTSAM itself is not installed here), and on a
GPU a C++ program that drives the shims the way OdometryEstimationGPU does reproduces the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
CPP = os.path.join(ROOT, "tests", "cpp")
GXX = "/usr/bin/g++"
GCC = "/usr/bin/gcc"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c_abi.c"
    src.write_text('#include "glim_b200.h"\nint main(void) { gb_linearized6 l; (void)l; return sizeof(gb_linearized6) == 122 * sizeof(double) ? 0 : 1; }\n')
    exe = tmp_path / "c_abi"
    subprocess.check_call([GCC, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{INC}", str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_shims_compile_standalone_and_gtsam_mode(tmp_path):
    subprocess.check_call([GXX, "-std=c++17", "-Wall", "-Wextra", "-Werror", f"-I{INC}", "-c", os.path.join(CPP, "shim_main.cpp"), "-o", str(tmp_path / "a.o")])
    subprocess.check_call([GXX, "-std=c++17", "-Wall", "-Wextra", "-Werror", f"-I{INC}", f"-I{os.path.join(CPP, 'gtsam_stub')}", "-c", os.path.join(CPP, "gtsam_mode_check.cpp"), "-o", str(tmp_path / "b.o")])


@pytest.mark.gpu
def test_cpp_shim_reproduces_oracle(tmp_path):
    from glim_b200 import capi, synth
    from oracle import oracle
    from tests import util

    exe = tmp_path / "shim_main"
    libdir = os.path.dirname(capi.SO_PATH)
    subprocess.check_call([GXX, "-std=c++17", "-O2", f"-I{INC}", os.path.join(CPP, "shim_main.cpp"), "-o", str(exe), f"-L{libdir}", "-lglim_b200", f"-Wl,-rpath,{libdir}"])
    pair = util.scan_pair()
    Tt = synth.pose(3.0, -1.0, 0.2, 0.4, 0.01, -0.02)
    T = synth.perturb(synth.inv_pose(pair["poses"][0]) @ pair["poses"][1], synth.rng_for(21), 0.01, 0.05)
    Ts = Tt @ T
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        np.array([len(pair["points"][0]), len(pair["points"][1])], np.int32).tofile(f)
        for k in (0, 1):
            np.ascontiguousarray(pair["points"][k]).tofile(f)
            util.cov_colmajor16(pair["covs"][k]).tofile(f)
        capi.pose16(Tt).tofile(f)
        capi.pose16(Ts).tofile(f)
    out = tmp_path / "out.bin"
    subprocess.check_call([str(exe), str(inp), str(out)])
    raw = np.fromfile(out, dtype=np.uint8)
    recs = np.frombuffer(raw[: 4 * 976].tobytes(), dtype=np.float64).reshape(4, 122)
    err, ov = np.frombuffer(raw[4 * 976 : 4 * 976 + 16].tobytes(), dtype=np.float64)
    nv = np.frombuffer(raw[4 * 976 + 16 : 4 * 976 + 32].tobytes(), dtype=np.int32)

    xyz0, cov0 = oracle.pack_cloud(pair["points"][0], util.cov_colmajor16(pair["covs"][0]))
    xyz1, cov1 = oracle.pack_cloud(pair["points"][1], util.cov_colmajor16(pair["covs"][1]))
    delta = synth.inv_pose(Tt) @ Ts
    for level, res in enumerate((0.25, 0.5)):
        m = oracle.GpuMap(xyz0, cov0, res)
        assert (nv[level], nv[2 + level]) == (m.num_voxels, m.num_buckets)
        ref = oracle.split122(oracle.linearize_gpumap(m, xyz1, cov1, delta)[0])
        for rec in (recs[level], recs[2 + level]):  # binary and unary forms see the same delta
            got = oracle.split122(rec)
            assert got["num_inliers"] == ref["num_inliers"] > 0
            for k in ("H_tt", "H_ss", "H_ts"):
                assert util.rel_err(got[k], ref[k]) < util.REL_TOL
            assert abs(got["error"] - ref["error"]) < util.REL_TOL * ref["error"]
        if level == 1:
            assert abs(err - ref["error"]) < util.REL_TOL * ref["error"]
            assert ov == oracle.overlap_gpumap([m], xyz1, [delta])
