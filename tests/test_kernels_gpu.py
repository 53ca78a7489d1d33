"""-m gpu: every CUDA kernel, through the C ABI, against a PyTorch fp32 reference of the same op."""
import pytest

from kernel_checks import CHECKS


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CHECKS))
def test_kernel(name):
    err, tol = CHECKS[name]()
    assert err == err and err <= tol, f"{name}: rel err {err:.3e} > tol {tol:.1e}"
