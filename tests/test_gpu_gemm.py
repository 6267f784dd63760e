"""-m gpu: tcgen05 GEMM kernels against the SIMT kernels and against fp32 torch.matmul."""
import numpy as np
import pytest
import torch

from tests._gpu import DEV

pytestmark = pytest.mark.gpu

PREC = {'fp32': 0, 'bf16': 1, 'bf16x3': 2}


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _ref(a, b, bias, relu, prec):
    if prec == 'bf16':
        a, b = _bf16(a), _bf16(b)
    c = a.double() @ b.double().T
    if bias is not None:
        c = c + bias.double()
    if relu:
        c = torch.relu(c)
    return c


@pytest.mark.parametrize('impl', [1, 0], ids=['simt', 'tc'])
@pytest.mark.parametrize('prec', ['fp32', 'bf16', 'bf16x3'])
@pytest.mark.parametrize('shape', [(384, 256, 256), (1000, 256, 128), (256, 144, 256), (512, 16, 256), (640, 256, 384),
                                   (20000, 256, 512)])
def test_gemm(prec, impl, shape):
    from refnerf_pl_b200 import ops
    m, n, k = shape
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(DEV)
    b = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV)
    c = ops.gemm_test(a, b, bias, True, PREC[prec], impl)
    ref = _ref(a, b, bias, True, prec)
    err = (c.double() - ref).abs().max().item()
    tol = {'fp32': 2e-5, 'bf16': 1e-4, 'bf16x3': 3e-4}[prec]   # bf16: exact products, fp32 accumulate
    assert err <= tol, (prec, impl, shape, err)


@pytest.mark.parametrize('impl', [1, 0], ids=['simt', 'tc'])
@pytest.mark.parametrize('prec', ['fp32', 'bf16', 'bf16x3'])
@pytest.mark.parametrize('shape', [(512, 128, 256), (4096, 128, 128), (3000, 16, 256), (40000, 128, 256),
                                   (3000, 256, 256), (20000, 144, 128), (70000, 256, 192)])
def test_wgrad(prec, impl, shape):
    from refnerf_pl_b200 import ops
    m, n, k = shape
    g = torch.Generator().manual_seed(m + n + k + 1)
    dy = torch.randn(m, n, generator=g).to(DEV)
    x = torch.randn(m, k, generator=g).to(DEV)
    c = ops.wgrad_test(dy, x, PREC[prec], impl)
    if prec == 'bf16':
        ref = _bf16(dy).double().T @ _bf16(x).double()
    else:
        ref = dy.double().T @ x.double()
    scale = np.sqrt(m)
    err = (c.double() - ref).abs().max().item() / scale
    tol = {'fp32': 1e-5, 'bf16': 5e-5, 'bf16x3': 2e-4}[prec]
    assert err <= tol, (prec, impl, shape, err)
