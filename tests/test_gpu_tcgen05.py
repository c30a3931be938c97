"""tcgen05 GEMM with fused bias + GELU epilogue (csrc/gemm_tcgen05.cu) against a plain PyTorch fp32 reference.
Variant 0 passed on B200 for every shape below except the last (added afterwards: the full GPT-2 c_fc shape, 768 CTAs
of the same per-tile kernel) and for the autograd test — log/gpu_tcgen05_test.log."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
# variant 1 (persistent CTAs, double-buffered TMEM) has been compiled and SASS-checked but not run yet: a protocol
# bug would trap the kernel and poison this process's CUDA context, so it only runs when asked for
VARIANTS = [0, 1] if os.environ.get("ADAPCC_EXPERIMENTAL", "0") == "1" else [0]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


@pytest.mark.parametrize("m,n,k,act", [(128, 256, 64, "none"), (256, 256, 128, "gelu"), (200, 384, 192, "gelu"),
                                       (1024, 3072, 768, "gelu"), (384, 128, 3072, "none"),
                                       (8192, 3072, 768, "gelu")])
@pytest.mark.parametrize("variant", VARIANTS)
def test_gemm_bias_act_matches_fp32_reference(dev, m, n, k, act, variant):
    from adapcc_b200.ops.gemm import linear_act

    torch.manual_seed(m + n + k)
    x = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) / k ** 0.5).bfloat16()
    b = torch.randn(n, device=dev).bfloat16()
    out, pre = linear_act(x, w, b, act, save_pre=True, variant=variant)
    torch.cuda.synchronize()
    u = x.float() @ w.float().t() + b.float()
    assert torch.allclose(pre.float(), u, atol=3e-2, rtol=2e-2), (pre.float() - u).abs().max()
    want = torch.nn.functional.gelu(u.bfloat16().float(), approximate="tanh") if act == "gelu" else u
    assert torch.allclose(out.float(), want, atol=3e-2, rtol=2e-2), (out.float() - want).abs().max()


def test_linear_gelu_autograd_matches_torch(dev):
    from adapcc_b200.ops.gemm import linear_gelu

    torch.manual_seed(3)
    x = torch.randn(4, 96, 256, device=dev).bfloat16().requires_grad_(True)
    w = (torch.randn(1024, 256, device=dev) / 16).bfloat16().requires_grad_(True)
    b = torch.randn(1024, device=dev).bfloat16().requires_grad_(True)
    dy = torch.randn(4, 96, 1024, device=dev).bfloat16()
    linear_gelu(x, w, b).backward(dy)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    torch.nn.functional.gelu(torch.nn.functional.linear(xr, wr, br), approximate="tanh").backward(dy.float())
    for got, want in ((x.grad, xr.grad), (w.grad, wr.grad), (b.grad, br.grad)):
        err = (got.float() - want).abs().max() / want.abs().max()
        assert err < 3e-2, err
