"""The C++ host-side mirror (include/aleo_b200.hpp): compiles against the product library; host logic
(layouts, EvaluationDomain::new, refusal without a device) on CPU, device calls under -m gpu."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(product_lib_path):
    exe = os.path.join(ROOT, "build", "test_mirror")
    libdir = os.path.dirname(product_lib_path)
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"),
                    "-L" + libdir, "-laleo_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True, cwd=ROOT)
    return exe


def test_cpp_mirror_host_logic(product_lib_path):
    out = subprocess.run([_build(product_lib_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "host logic ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_on_device(product_lib_path):
    out = subprocess.run([_build(product_lib_path), "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "gpu mirror ok" in out.stdout, out.stdout + out.stderr
