"""CPU suite: the fp64 two-view solvers the CUDA verifier compiles (gtsfm_b200/csrc/ransac_math.cuh) built for the host."""
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_fivept_svd_pose_host_build(tmp_path):
    exe = tmp_path / "test_ransac_math"
    cxx = shutil.which("g++")
    assert cxx, "g++ is required"
    subprocess.run([cxx, "-O2", "-std=c++17", "-x", "c++", str(ROOT / "tests/cpp/test_ransac_math.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
