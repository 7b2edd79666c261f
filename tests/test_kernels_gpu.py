"""pytest -m gpu: every entry point of libopp_b200.so against fp64/fp32 torch math on the same
inputs (tests/kernel_checks.py), in both operand modes (split = fp16x3 parity mode, plain fp16)."""
import pytest

from tests import kernel_checks


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(kernel_checks.CHECKS))
def test_kernel(name):
    kernel_checks.CHECKS[name]()
