"""tcgen05 building blocks: 128 x N x K GEMM with the 3xFP16 split against fp64."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(N, K, variant=0, seed=0):
    from mopoe_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(128, K, generator=g) * 2).cuda()
    B = (torch.randn(N, K, generator=g) * 0.1).cuda()
    D = torch.zeros(128, N, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    T = _lib.selftest_lib()                  # test-only library, not part of libmopoe_b200.so
    rc = T.mopoe_umma_selftest(p(A), p(B), p(D), N, K, variant, p(err), None)
    assert rc == 0, T.mopoe_last_error().decode()
    torch.cuda.synchronize()
    want = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T).max()
    return float((D.double() - want).abs().max() / scale), int(err.item())


@pytest.mark.parametrize("N,K", [(48, 256), (224, 48), (16, 16), (256, 64), (64, 128)])
def test_umma_split_gemm(N, K):
    rel, err = _run(N, K)
    assert err == 0, "mbarrier wait timed out"
    assert rel <= 2e-6, rel


@pytest.mark.parametrize("N,K", [(48, 256), (96, 64), (16, 32)])
def test_umma_split_gemm_a_from_tmem(N, K):
    """A operand written to TMEM with tcgen05.st (thread = row) and consumed by tcgen05.mma [d], [a], b."""
    rel, err = _run(N, K, variant=2)
    assert err == 0, "mbarrier wait timed out"
    assert rel <= 2e-6, rel


@pytest.mark.parametrize("N,K", [(16, 128), (48, 64), (16, 16)])
def test_umma_split_gemm_mn_major_operands(N, K):
    """Both operands MN-major (instruction-descriptor bits 15/16; LBO = stride between 8-element K groups,
    SBO = stride between 8-element MN chunks): the layout a K-major tile has when it is read transposed."""
    rel, err = _run(N, K, variant=3)
    assert err == 0, "mbarrier wait timed out"
    assert rel <= 2e-6, rel
