"""The tcgen05/TMEM/TMA GEMM kernel alone, against a float64 numpy product (vt_debug_gemm entry of the C ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(built):
    from gstreamer_vit_tracker_b200 import api as _api
    return _api


def _rand(shape, seed):
    from gstreamer_vit_tracker_b200.synth import hash_u64
    n = int(np.prod(shape))
    u = (hash_u64(seed, n) >> np.uint64(11)).astype(np.float64) / (1 << 53)
    return ((u * 2 - 1)).astype(np.float32).reshape(shape)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (320, 192, 192), (320, 576, 192), (320, 192, 768), (256, 192, 768), (5120, 768, 192), (100, 64, 128), (64, 64, 768)])
@pytest.mark.parametrize("nsplit", [1, 2, 3], ids=["bf16", "fp16", "bf16x3"])
def test_gemm_matches_float64(api, M, N, K, nsplit):
    A, W, b = _rand((M, K), 1 + M), _rand((N, K), 2 + N) * (1.0 / np.sqrt(K)), _rand((N,), 3)
    C, err = api.debug_gemm(A, W, b, nsplit=nsplit)
    assert err == 0, "a bounded mbarrier wait expired"
    ref = A.astype(np.float64) @ W.astype(np.float64).T + b
    scale = float(np.abs(ref).max())
    e = float(np.abs(C - ref).max()) / scale
    tol = {3: 2e-5, 2: 1.5e-3, 1: 1e-2}[nsplit]
    assert e < tol, (M, N, K, nsplit, e)
    if nsplit == 2:  # must equal the fp16-rounded-operand product (operand rounding alone is ~5e-4)
        Ah, Wh = A.astype(np.float16).astype(np.float64), W.astype(np.float16).astype(np.float64)
        assert float(np.abs(C - (Ah @ Wh.T + b)).max()) / scale < 5e-5
    if nsplit == 1:  # must equal the bf16-rounded-operand product (proves the operands really are bf16 and K is fully reduced)
        import torch
        Ab = torch.from_numpy(A).bfloat16().double().numpy()
        Wb = torch.from_numpy(W).bfloat16().double().numpy()
        ref1 = Ab @ Wb.T + b
        assert float(np.abs(C - ref1).max()) / scale < 2e-5


def test_gemm_gelu_epilogue(api):
    from math import erf
    A, W = _rand((320, 192), 7), _rand((768, 192), 8) * 0.1
    C, err = api.debug_gemm(A, W, None, nsplit=3, gelu=True)
    assert err == 0
    z = A.astype(np.float64) @ W.astype(np.float64).T
    ref = 0.5 * z * (1 + np.vectorize(erf)(z / np.sqrt(2)))
    assert float(np.abs(C - ref).max()) < 3e-5
