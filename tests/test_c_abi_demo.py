"""The drop-in boundary used from plain C (examples/c_abi_demo.c): include/unet_b200.h + libunet_b200.so, no Python / torch in
the process. CPU: compiles with gcc -std=c99 and runs the host-side part (plan geometry, argument validation, the loud
"no sm_100 device" stop). GPU: the same binary builds the default network with seeded weights, runs frames through
unet_b200_infer_u8_host_stream and checks mask == (prob > 0.5) * 255 and run-to-run identity itself."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unet-lane-detection_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_demo")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", PKG, "-lunet_b200", "-L", os.path.join(CUDA, "lib64"),
           "-lcudart", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_host_side(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU run of the same program")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "18 convs, 23 layers, 22 kernels per pass" in r.stdout
    assert "features[1]=0 must be in [1,4096]" in r.stdout
    assert "host-side checks only" in r.stdout and "OK" not in r.stdout.splitlines()[-1]


@pytest.mark.gpu
def test_c_program_runs_the_network(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.splitlines()[-1] == "OK" and "second run bit-identical" in r.stdout
