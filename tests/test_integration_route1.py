"""INTEGRATION.md route 1 (link the reference's unmodified bindings against csrc/): nanobind is not installed here, so a
stand-in translation unit that uses exactly the names and call shapes of the reference's src/main.cpp
(tests/route1/stub_bindings.cpp) is compiled TOGETHER with the product sources, linked and run on the CPU branch."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SS = os.path.join(os.environ.get("CHOLESPY_REF", "/root/reference"), "ext", "suitesparse-metis-for-windows", "SuiteSparse")


def test_stub_bindings_compile_link_and_run(tmp_path, native_lib):
    if not os.path.isdir(SS):
        pytest.skip("CHOLMOD headers (reference checkout) absent")
    csrc = os.path.join(ROOT, "cholespy_b200", "csrc")
    lib = os.path.join(ROOT, "cholespy_b200", "lib")
    image = os.path.join(lib, "_gen", "kernels_image.cpp")
    assert os.path.exists(image), "build the library first"
    sources = [os.path.join(csrc, f) for f in ("cholesky_solver.cpp", "gpu_solver.cpp", "cuda_driver.cpp", "factor_host.cpp",
                                               "panel_schedule.cpp", "front_schedule.cpp", "disk_cache.cpp")]
    exe = str(tmp_path / "route1_stub")
    cmd = ["g++", "-O1", "-std=c++17", "-w", "-I", csrc, "-I", os.path.join(SS, "CHOLMOD", "Include"),
           "-I", os.path.join(SS, "SuiteSparse_config"), "-o", exe, os.path.join(ROOT, "tests", "route1", "stub_bindings.cpp")] + sources + [
           image, "-L", lib, "-lcholmod_b200", "-ldl", "-lpthread", "-Wl,-rpath," + lib]
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "CholeskySolver<float> residual" in out and "CholeskySolver<double> residual" in out
    assert out.count("raised: 1") == 2
