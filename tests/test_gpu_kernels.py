"""GPU: single-kernel parity of the tcgen05 tap-GEMM / wgrad kernels through the C ABI, against torch fp32 on the
same bf16-rounded operands (a floating-point kernel: tolerance 1.5e-2 of the reference max, bf16 output rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _cases():
    import kernel_cases
    return list(kernel_cases.CASES)


@pytest.mark.parametrize("name", _cases())
def test_kernel_case(name):
    import kernel_cases
    err, ref = kernel_cases.CASES[name]()
    assert err == err, "NaN in output"
    assert err <= 1.5e-2 * max(ref, 1e-6), f"{name}: max abs err {err} vs ref max {ref}"
