"""Fused residual-add + LayerNorm (csrc/layer_epilogue.cu) against torch's add + F.layer_norm (M2F:1049-1050)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    from weed_instance_segmentation_b200 import build
    build.build()


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows", [1, 7, 64, 1000, 21504])
@pytest.mark.parametrize("C", [128, 256, 512])
@pytest.mark.parametrize("xdt,rdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                     (torch.bfloat16, torch.bfloat16)])
def test_add_layer_norm_matches_torch(rows, C, xdt, rdt):
    from weed_instance_segmentation_b200.layer_norm import add_layer_norm
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + C)
    x = torch.randn(2, rows, C, device="cuda", generator=g).to(xdt).requires_grad_(True)
    r = (3.0 * torch.randn(2, rows, C, device="cuda", generator=g) + 0.5).to(rdt).requires_grad_(True)
    w = (1.0 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    go = torch.randn(2, rows, C, device="cuda", generator=g)
    y = add_layer_norm(x, r, w, b, 1e-5)
    assert y.dtype == torch.float32
    y.backward(go)
    got = [y.detach(), x.grad.clone(), r.grad.clone(), w.grad.clone(), b.grad.clone()]
    # reference in fp64 on the same (rounded) inputs
    xr, rr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, r, w, b))
    yr = F.layer_norm(rr + xr, (C,), wr, br, 1e-5)
    yr.backward(go.double())
    want = [yr.detach(), xr.grad, rr.grad, wr.grad, br.grad]
    bars = [1e-5, 1e-5 if xdt == torch.float32 else 1e-2, 1e-5 if rdt == torch.float32 else 1e-2, 2e-5, 2e-5]
    for name, a, c, bar in zip(("y", "grad_x", "grad_r", "grad_w", "grad_b"), got, want, bars):
        assert a.shape == c.shape
        assert _rel(a, c) <= bar, (name, _rel(a, c))


def test_add_layer_norm_errors():
    from weed_instance_segmentation_b200 import MSDAError
    from weed_instance_segmentation_b200.layer_norm import add_layer_norm
    x = torch.zeros(4, 96, device="cuda")
    with pytest.raises(MSDAError):  # 96 channels: not a multiple of 128
        add_layer_norm(x, x, torch.ones(96, device="cuda"), torch.zeros(96, device="cuda"))
    with pytest.raises(ValueError):
        add_layer_norm(torch.zeros(4, 128, device="cuda"), torch.zeros(5, 128, device="cuda"),
                       torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"))
    with pytest.raises(RuntimeError):
        add_layer_norm(torch.zeros(4, 128), torch.zeros(4, 128), torch.ones(128), torch.zeros(128))
    y = add_layer_norm(torch.zeros(0, 128, device="cuda"), torch.zeros(0, 128, device="cuda"),
                       torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"))
    assert y.shape == (0, 128)


@pytest.mark.parametrize("clamp", [0.75, float(torch.finfo(torch.float32).max - 1000)])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_add_layer_norm_with_clamp_matches_torch(clamp, xdt):
    """The encoder layer's closing clamp (M2F:1062-1065) folded into the fused kernels: same values and the same
    gradients as torch.clamp(F.layer_norm(...)) -- with a small bound (many elements clamped, their gradient masked), at
    the layer's real bound (identity), and with Inf / NaN planted in the input (NaN stays NaN, no gradient through it)."""
    from weed_instance_segmentation_b200.layer_norm import add_layer_norm
    C, rows = 256, 333
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(rows, C, device="cuda", generator=g).to(xdt).requires_grad_(True)
    r = (2.0 * torch.randn(rows, C, device="cuda", generator=g)).requires_grad_(True)
    w = (1.0 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    go = torch.randn(rows, C, device="cuda", generator=g)
    y = add_layer_norm(x, r, w, b, 1e-5, clamp=clamp)
    y.backward(go)
    xr, rr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, r, w, b))
    yr = torch.clamp(F.layer_norm(rr + xr, (C,), wr, br, 1e-5), min=-clamp, max=clamp)
    yr.backward(go.double())
    if clamp < 1:
        assert 0.2 < (yr.detach().abs() >= clamp).double().mean().item() < 0.8  # the bound is active
        # elements within round-off of the bound may fall on either side in fp32 vs fp64: leave them out of the gradient check
        pre = F.layer_norm(rr.detach() + xr.detach(), (C,), wr.detach(), br.detach(), 1e-5)
        safe = ((pre.abs() - clamp).abs() > 1e-4).all(dim=-1, keepdim=True).double()  # whole rows (LayerNorm couples a row)
        assert safe.mean().item() > 0.3
    else:
        safe = torch.ones(rows, 1, device="cuda", dtype=torch.float64)
    bar_in = 1e-5 if xdt == torch.float32 else 1e-2
    assert _rel(y.detach(), yr.detach()) <= 1e-5
    assert _rel(x.grad * safe, xr.grad * safe) <= bar_in and _rel(r.grad * safe, rr.grad * safe) <= 1e-5
    if clamp > 1:
        assert _rel(w.grad, wr.grad) <= 2e-5 and _rel(b.grad, br.grad) <= 2e-5
    # non-finite activations: same treatment as torch.clamp
    x2 = torch.randn(8, C, device="cuda", generator=g)
    x2[1, 3], x2[2, 5], x2[4, 7] = float("inf"), float("nan"), -float("inf")
    r2 = torch.zeros_like(x2)
    big = float(torch.finfo(torch.float32).max - 1000)
    got = add_layer_norm(x2, r2, w.detach(), b.detach(), 1e-5, clamp=big)
    want = torch.clamp(F.layer_norm(x2, (C,), w.detach(), b.detach(), 1e-5), min=-big, max=big)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    ok = ~torch.isnan(want)
    assert torch.allclose(got[ok], want[ok], rtol=1e-5, atol=1e-5)
