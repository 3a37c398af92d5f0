"""Projection with the fused bias-gradient reduction (linear.py, msda_b200_column_sum) against F.linear."""
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


@pytest.mark.parametrize("rows", [1, 33, 5000, 172032])
@pytest.mark.parametrize("cols", [8, 96, 192, 256, 1024])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_column_sum(rows, cols, dtype):
    from weed_instance_segmentation_b200.linear import column_sum
    if rows * cols > 60_000_000:
        rows = 60_000_000 // cols
    m = torch.randn(rows, cols, device="cuda").to(dtype)
    got = column_sum(m)
    want = m.double().sum(0)
    assert got.dtype == torch.float32 and got.shape == (cols,)
    assert ((got.double() - want).abs().max() / want.abs().max().clamp_min(1.0)).item() < 1e-5


@pytest.mark.parametrize("autocast", [False, True])
def test_linear_matches_torch(autocast):
    from weed_instance_segmentation_b200.linear import linear
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(4, 1000, 256, device="cuda", generator=g)
    w = (0.05 * torch.randn(192, 256, device="cuda", generator=g))
    b = torch.randn(192, device="cuda", generator=g)
    go = torch.randn(4, 1000, 192, device="cuda", generator=g)
    res = {}
    for name, fn in (("ref", F.linear), ("new", linear)):
        xs, ws, bs = (t.clone().requires_grad_(True) for t in (x, w, b))
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = fn(xs, ws, bs)
        y.backward(go.to(y.dtype))
        res[name] = (y.detach().float(), xs.grad, ws.grad, bs.grad)
    bar = 2e-2 if autocast else 1e-5
    for a, c, nm in zip(res["new"], res["ref"], ("y", "grad_x", "grad_w", "grad_b")):
        assert a.dtype == c.dtype and a.shape == c.shape
        assert _rel(a, c) <= bar, (nm, _rel(a, c))
    # the bias gradient itself is more accurate than torch's bf16 reduction: compare with fp64
    gb64 = go.to(torch.bfloat16 if autocast else torch.float32).double().reshape(-1, 192).sum(0)
    assert _rel(res["new"][3], gb64) <= (1e-5 if not autocast else 1e-5)


def test_column_sum_errors():
    from weed_instance_segmentation_b200 import MSDAError
    from weed_instance_segmentation_b200.linear import column_sum
    with pytest.raises(MSDAError):
        column_sum(torch.zeros(4, 12, device="cuda", dtype=torch.bfloat16))  # 12 columns: not a multiple of 8
    assert column_sum(torch.zeros(0, 16, device="cuda")).abs().sum().item() == 0


def test_query_value_cast_is_the_stock_sequence():
    """bfloat16(hidden + pos) and bfloat16(hidden) from one kernel: bit-identical to the fp32 add + two casts of the
    stock module under autocast (M2F:936-937, 947), and so are the gradients (two casts + an add)."""
    from weed_instance_segmentation_b200.linear import query_value_cast
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape in ((2, 333, 256), (1, 8), (3, 0, 256)):
        h = torch.randn(shape, device="cuda", generator=g).requires_grad_(True)
        p = (0.3 * torch.randn(shape, device="cuda", generator=g)).requires_grad_(True)
        q, v = query_value_cast(h, p)
        assert q.dtype == v.dtype == torch.bfloat16
        assert torch.equal(q, (h + p).to(torch.bfloat16)) and torch.equal(v, h.to(torch.bfloat16))
        gq = torch.randn(shape, device="cuda", generator=g).to(torch.bfloat16)
        gv = torch.randn(shape, device="cuda", generator=g).to(torch.bfloat16)
        torch.autograd.backward([q, v], [gq, gv])
        assert torch.equal(h.grad, gq.float() + gv.float()) and torch.equal(p.grad, gq.float())
        # only one of the two operands used downstream; pos without gradient
        h2 = h.detach().clone().requires_grad_(True)
        q2, v2 = query_value_cast(h2, p.detach())
        v2.backward(gv)
        assert torch.equal(h2.grad, gv.float())
    with pytest.raises(ValueError):
        query_value_cast(torch.zeros(4, 8, device="cuda"), torch.zeros(4, 16, device="cuda"))
    with pytest.raises(TypeError):
        query_value_cast(torch.zeros(4, 8, device="cuda", dtype=torch.bfloat16), torch.zeros(4, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        query_value_cast(torch.zeros(4, 8), torch.zeros(4, 8))


@pytest.mark.parametrize("rows,out_f,in_f", [(1, 256, 256), (777, 1024, 256), (5000, 256, 1024), (21504, 192, 256), (33, 96, 256)])
def test_f32_projection_on_tensor_cores_is_at_least_as_accurate_as_sgemm(rows, out_f, in_f, monkeypatch):
    """float32 projections through cuBLASLt's emulated-float32 compute type (csrc/gemm_f32.cu): forward (bias, bias + ReLU)
    and both backward products against float64, and never worse than twice the error of torch's own SGEMM path."""
    from weed_instance_segmentation_b200 import linear as L
    if not L.f32_gemm_available():
        pytest.skip("the CUDA toolkit's cuBLASLt with float32 emulation is not available on this box")
    g = torch.Generator(device="cuda").manual_seed(rows + out_f)
    x = (2.0 * torch.randn(3, rows, in_f, device="cuda", generator=g)).requires_grad_(True)
    w = (0.1 * torch.randn(out_f, in_f, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(out_f, device="cuda", generator=g)).requires_grad_(True)
    go = torch.randn(3, rows, out_f, device="cuda", generator=g)

    def run(fn):
        for t in (x, w, b):
            t.grad = None
        y = fn(x, w, b)
        y.backward(go)
        return [y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone()]

    xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    for relu in (False, True):
        fn = L.linear_relu if relu else L.linear
        got = run(fn)
        monkeypatch.setattr(L, "_F32_EMULATED", False)
        native = run(fn)  # torch's SGEMM through the same autograd functions
        monkeypatch.setattr(L, "_F32_EMULATED", True)
        for t in (xr, wr, br):
            t.grad = None
        yr = F.linear(xr, wr, br)
        yr = F.relu(yr) if relu else yr
        yr.backward(go.double())
        want = [yr.detach(), xr.grad, wr.grad, br.grad]
        for name, a, n, c in zip(("y", "grad_x", "grad_w", "grad_b"), got, native, want):
            assert a.dtype == torch.float32 and a.shape == c.shape
            if relu and name != "y":
                continue  # a pre-activation within round-off of 0 may switch its mask between the two float32 paths
            e, en = _rel(a, c), _rel(n, c)
            assert e <= 2e-6, (relu, name, e)
            assert e <= 2 * en + 1e-7, (relu, name, e, en)
        if relu:  # gradients: compare where both float32 paths agree on the mask
            same = ((got[0] > 0) == (native[0] > 0)).all()
            if same:
                for name, a, c in zip(("grad_x", "grad_w", "grad_b"), got[1:], want[1:]):
                    assert _rel(a, c) <= 5e-6, (name, _rel(a, c))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,cols", [(1, 8), (37, 1024), (5000, 256), (0, 64)])
def test_relu_backward_column_sum_matches_torch(rows, cols, dtype):
    """ReLU backward mask and the bias gradient from one kernel: the masked gradient is bit-identical to aten
    threshold_backward, the column sum matches its float64 sum."""
    from weed_instance_segmentation_b200.linear import relu_backward_column_sum
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    y = torch.relu(torch.randn(rows, cols, device="cuda", generator=g)).to(dtype)
    gy = torch.randn(rows, cols, device="cuda", generator=g).to(dtype)
    got, col = relu_backward_column_sum(gy, y)
    want = torch.ops.aten.threshold_backward(gy, y, 0)
    assert got.dtype == dtype and torch.equal(got, want)
    ref = want.double().sum(0)
    assert col.dtype == torch.float32 and col.shape == (cols,)
    assert (col.double() - ref).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1.0) + 1e-6 * max(rows, 1) ** 0.5
