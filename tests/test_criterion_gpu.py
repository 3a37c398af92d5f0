"""point_sample kernel and the batched criterion on a B200 against the reference's own code on the same device.

Reference: ``sample_point`` (M2F:245-274, i.e. grid_sample) for the kernel, ``Mask2FormerLoss`` (M2F:476-793) with
the same CUDA generator state for the criterion. Floating point: 1e-5 relative on sampled values / losses.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from test_criterion_host import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ps():
    from weed_instance_segmentation_b200 import _cabi, build
    from weed_instance_segmentation_b200.point_sample import point_sample
    build.build()
    _cabi.load()
    return point_sample


def _reference_rows(sources, src_id, plane_id, coords, coord_row):
    rows = []
    for s, p, c in zip(src_id, plane_id, coord_row):
        plane = sources[int(s)][int(p)][None, None].float()
        pts = coords[int(c)][None, :, None, :]
        rows.append(F.grid_sample(plane, 2.0 * pts - 1.0, align_corners=False)[0, 0, :, 0])
    return torch.stack(rows)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_point_sample_matches_grid_sample(ps, dtype):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(5, 17, 23, device="cuda", generator=g).to(dtype).requires_grad_(True)
    b = torch.randn(3, 64, 40, device="cuda", generator=g).to(dtype).requires_grad_(True)
    c = (torch.rand(2, 9, 9, device="cuda", generator=g) > 0.5).float()
    coords = torch.rand(4, 301, 2, device="cuda", generator=g) * 1.3 - 0.15       # some points outside
    coords[0, :3] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.5, float("nan")]], device="cuda")
    src = np.array([0, 0, 1, 2, 1, 0, 2, 1])
    plane = np.array([4, 0, 2, 1, 0, 4, 0, 2])
    crow = np.array([0, 1, 2, 3, 0, 2, 3, 2])
    out = ps([a, b, c], src, plane, coords, crow)
    go = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(go)
    a2, b2 = a.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    safe = coords.clone()
    safe[0, 2] = 2.0  # grid_sample propagates NaN; the kernel treats it as "outside" (zero), like the MSDA op
    want = _reference_rows([a2, b2, c], src, plane, safe, crow)
    want.backward(go)
    assert rel_err(out.detach().cpu().numpy(), want.detach().cpu().numpy()) <= 1e-5
    tol = 1e-5 if dtype == torch.float32 else 8e-3   # bf16: the gradient is rounded once on the way out
    assert rel_err(a.grad.float().cpu().numpy(), a2.grad.cpu().numpy()) <= tol
    assert rel_err(b.grad.float().cpu().numpy(), b2.grad.cpu().numpy()) <= tol
    assert a.grad.dtype == dtype


def test_point_sample_errors_and_empty(ps):
    a = torch.zeros(2, 4, 4, device="cuda")
    coords = torch.rand(1, 8, 2, device="cuda")
    assert ps([a], [], [], coords, []).shape == (0, 8)
    with pytest.raises(ValueError):
        ps([a], [0], [2], coords, [0])
    with pytest.raises(ValueError):
        ps([a], [0], [0], coords, [1])
    with pytest.raises(RuntimeError):
        ps([a.cpu()], [0], [0], coords, [0])
    with pytest.raises(TypeError):
        ps([a.half()], [0], [0], coords, [0])


def _run(loss, masks, classes, mask_labels, class_labels, seed):
    masks = [m.clone().requires_grad_(True) for m in masks]
    classes = [c.clone().requires_grad_(True) for c in classes]
    aux = [{"masks_queries_logits": m, "class_queries_logits": c} for m, c in zip(masks[:-1], classes[:-1])]
    torch.manual_seed(seed)
    out = loss(masks[-1], classes[-1], mask_labels, class_labels, aux if aux else None)
    total = sum(v * (i + 1) for i, v in enumerate(out.values()))
    total.backward()
    return out, [m.grad for m in masks], [c.grad for c in classes]


@pytest.mark.parametrize("n_tgt", [(2, 5, 1), (4, 0, 3), (14, 1, 2)])
def test_criterion_matches_reference_on_gpu(ps, n_tgt):
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(0, n_tgt=n_tgt, num_points=200)
    loss = loss.cuda()
    masks, classes = [m.cuda() for m in masks], [c.cuda() for c in classes]
    mask_labels, class_labels = [m.cuda() for m in mask_labels], [c.cuda() for c in class_labels]
    want, gm_want, gc_want = _run(loss, masks, classes, mask_labels, class_labels, seed=7)
    mine = convert_criterion(copy.deepcopy(loss))
    got, gm_got, gc_got = _run(mine, masks, classes, mask_labels, class_labels, seed=7)
    assert list(got) == list(want)
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=2e-5, atol=1e-6), (k, float(got[k]), float(want[k]))
    for a, b in zip(gm_got + gc_got, gm_want + gc_want):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4


def test_criterion_full_size_matches_reference(ps):
    """BASELINE config 4 geometry of the loss: 10 layers, batch 8, 100 queries, 256x256 logits, 1024x1024 targets."""
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(
        1, B=8, Q=100, C=5, L=10, h=256, w=256, H=1024, W=1024, n_tgt=(3, 9, 14, 1, 20, 7, 5, 11), num_points=12544)
    loss = loss.cuda()
    masks, classes = [m.cuda() for m in masks], [c.cuda() for c in classes]
    mask_labels, class_labels = [m.cuda() for m in mask_labels], [c.cuda() for c in class_labels]
    want, gm_want, _ = _run(loss, masks, classes, mask_labels, class_labels, seed=3)
    mine = convert_criterion(copy.deepcopy(loss))
    got, gm_got, _ = _run(mine, masks, classes, mask_labels, class_labels, seed=3)
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-4, atol=1e-6), (k, float(got[k]), float(want[k]))
    for a, b in zip(gm_got, gm_want):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-3


def test_uint8_and_bool_target_masks_give_the_same_losses(ps):
    """Binary target masks may stay one byte per pixel (synth.collate_batch(mask_dtype=torch.uint8))."""
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(2, num_points=200)
    mine = convert_criterion(loss.cuda())
    masks, classes = [m.cuda() for m in masks], [c.cuda() for c in classes]
    class_labels = [c.cuda() for c in class_labels]
    outs = []
    for dt in (torch.float32, torch.uint8, torch.bool):
        labels = [m.to(dt).cuda() for m in mask_labels]
        out, gm, _ = _run(mine, masks, classes, labels, class_labels, seed=5)
        outs.append((out, gm))
    for out, gm in outs[1:]:
        for k in outs[0][0]:
            assert torch.equal(out[k], outs[0][0][k]), k
