"""The C++ drop-in classes (cxx/) over the C ABI: compile check on CPU, full run on the GPU."""
import os
import subprocess

import pytest


def _bin():
    from nonlinear_optimizer_for_slam_b200 import build as nlo_build
    return nlo_build.build_cxx_example()


def test_cxx_example_compiles_against_the_abi_without_cuda_headers():
    path = _bin()
    assert os.path.exists(path)
    # the reference-facing headers must not need CUDA or torch: g++ alone built them
    out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "libnlo_cuda.so" in out
    assert "libtorch" not in out


def test_cxx_example_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = subprocess.run([_bin()], capture_output=True, text=True)
    assert res.returncode != 0
    assert "no usable sm_100 GPU" in res.stderr


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_cxx_dropin_known_answers():
    """PnP known answer (COST 2.33228e-11, iter 6), NDT 6-DoF and 3-DoF registrations through
    MahalanobisDistanceMinimizerCuda / ...Cuda3DOF / ReprojectionErrorMinimizerCuda."""
    res = subprocess.run([_bin()], capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stderr[-4000:]
    assert "ALL CXX TESTS PASSED" in res.stderr
    assert "COST: 2.33228e-11, iter: 6" in res.stderr
