"""Host logic of the batched loss / matcher path (criterion.py) against the reference criterion on the CPU.

The reference here is HF's own ``Mask2FormerLoss`` (M2F:476-793) run with the same generator state; the sampler is
injected (a plain grid_sample loop, i.e. the reference's ``sample_point``), so this checks the batching, the random
number order, the assignment bookkeeping and the loss formulas -- not the CUDA kernel (tests/test_criterion_gpu.py).
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def grid_sample_sampler(sources, src_id, plane_id, coords, coord_row):
    rows = []
    for s, p, c in zip(src_id, plane_id, coord_row):
        plane = sources[int(s)][int(p)][None, None].float()
        pts = coords[int(c)][None, :, None, :]
        rows.append(F.grid_sample(plane, 2.0 * pts - 1.0, align_corners=False)[0, 0, :, 0])
    return torch.stack(rows) if rows else torch.zeros(0, coords.shape[1])


def make_problem(seed, B=3, Q=12, C=3, L=3, h=16, w=20, H=64, W=80, n_tgt=(2, 5, 1), num_points=48):
    from transformers import Mask2FormerConfig
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerLoss
    cfg = Mask2FormerConfig(num_labels=C, num_queries=Q, train_num_points=num_points)
    weight_dict = {"loss_cross_entropy": cfg.class_weight, "loss_mask": cfg.mask_weight, "loss_dice": cfg.dice_weight}
    loss = Mask2FormerLoss(cfg, weight_dict)
    g = torch.Generator().manual_seed(seed)
    masks = [torch.randn(B, Q, h, w, generator=g) * 3 for _ in range(L)]
    classes = [torch.randn(B, Q, C + 1, generator=g) for _ in range(L)]
    mask_labels = [(torch.rand(n, H, W, generator=g) > 0.6).float() for n in n_tgt]
    class_labels = [torch.randint(0, C, (n,), generator=g) for n in n_tgt]
    return loss, masks, classes, mask_labels, class_labels


def run(loss, masks, classes, mask_labels, class_labels, seed):
    masks = [m.clone().requires_grad_(True) for m in masks]
    classes = [c.clone().requires_grad_(True) for c in classes]
    aux = [{"masks_queries_logits": m, "class_queries_logits": c} for m, c in zip(masks[:-1], classes[:-1])]
    torch.manual_seed(seed)
    out = loss(masks[-1], classes[-1], mask_labels, class_labels, aux if aux else None)
    total = sum(v * (i + 1) for i, v in enumerate(out.values()))
    total.backward()
    return out, [m.grad for m in masks], [c.grad for c in classes]


@pytest.mark.parametrize("n_tgt", [(2, 5, 1), (4, 0, 3), (14, 1, 2)])
def test_batched_criterion_matches_reference(n_tgt):
    from weed_instance_segmentation_b200.criterion import convert_criterion, restore_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(0, n_tgt=n_tgt)
    want, gm_want, gc_want = run(loss, masks, classes, mask_labels, class_labels, seed=7)
    mine = convert_criterion(copy.deepcopy(loss), sampler=grid_sample_sampler)
    got, gm_got, gc_got = run(mine, masks, classes, mask_labels, class_labels, seed=7)
    assert list(got) == list(want)
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-6), (k, float(got[k]), float(want[k]))
    for a, b in zip(gm_got + gc_got, gm_want + gc_want):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7), float((a - b).abs().max())
    # and back
    restore_criterion(mine)
    again, _, _ = run(mine, masks, classes, mask_labels, class_labels, seed=7)
    for k in want:
        assert torch.equal(again[k], want[k])


def test_assignments_match_reference_matcher():
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(3, L=1)
    torch.manual_seed(11)
    want = loss.matcher(masks[0], classes[0], mask_labels, class_labels)
    mine = convert_criterion(copy.deepcopy(loss), sampler=grid_sample_sampler)
    torch.manual_seed(11)
    with torch.no_grad():
        mine(masks[0], classes[0], mask_labels, class_labels, None)
    for (wi, wj), (gi, gj) in zip(want, mine.last_indices):
        assert np.array_equal(wi.numpy(), np.asarray(gi)) and np.array_equal(wj.numpy(), np.asarray(gj))


def test_state_dict_and_class_are_drop_in():
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerLoss
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, *_ = make_problem(0)
    keys = list(loss.state_dict())
    mine = convert_criterion(loss)
    assert isinstance(mine, Mask2FormerLoss) and mine is loss
    assert list(mine.state_dict()) == keys


def test_default_sampler_refuses_cpu_tensors():
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(0, L=1)
    mine = convert_criterion(loss)
    with pytest.raises(RuntimeError, match="CUDA"):
        mine(masks[0], classes[0], mask_labels, class_labels, None)


def test_converted_loss_survives_pickle_and_deepcopy():
    import pickle

    from weed_instance_segmentation_b200 import criterion
    loss, *_ = make_problem(0)
    mine = criterion.convert_criterion(loss)
    again = pickle.loads(pickle.dumps(mine))
    assert type(again) is criterion.B200Mask2FormerLoss and type(copy.deepcopy(mine)) is type(mine)
    assert torch.equal(again.empty_weight, mine.empty_weight)


def _compare(loss, masks, classes, mask_labels, class_labels, seed=5):
    from weed_instance_segmentation_b200.criterion import convert_criterion
    want, gm_want, gc_want = run(loss, masks, classes, mask_labels, class_labels, seed=seed)
    mine = convert_criterion(copy.deepcopy(loss), sampler=grid_sample_sampler)
    got, gm_got, gc_got = run(mine, masks, classes, mask_labels, class_labels, seed=seed)
    assert list(got) == list(want)
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-6), (k, float(got[k]), float(want[k]))
    for a, b in zip(gm_got + gc_got, gm_want + gc_want):
        a = torch.zeros_like(b) if a is None else a  # the injected sampler returns a constant for zero rows
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7), float((a - b).abs().max())


def test_single_layer_single_image():
    """No auxiliary predictions (the inference-time loss call shape) and a batch of one."""
    loss, masks, classes, mask_labels, class_labels = make_problem(4, B=1, L=1, n_tgt=(3,))
    _compare(loss, masks, classes, mask_labels, class_labels)


def test_all_points_by_importance_and_no_oversampling():
    """importance_sample_ratio = 1 (no uniformly random tail: the reference then skips one torch.rand per layer) and
    oversample_ratio = 1 (top-k over exactly num_points candidates)."""
    loss, masks, classes, mask_labels, class_labels = make_problem(5, L=2)
    loss.importance_sample_ratio = 1.0
    _compare(loss, masks, classes, mask_labels, class_labels)
    loss.importance_sample_ratio = 0.75
    loss.oversample_ratio = 1.0
    _compare(loss, masks, classes, mask_labels, class_labels)


def test_non_default_cost_weights_change_the_assignment_consistently():
    loss, masks, classes, mask_labels, class_labels = make_problem(6, L=2, n_tgt=(6, 6, 6))
    loss.matcher.cost_class, loss.matcher.cost_mask, loss.matcher.cost_dice = 5.0, 0.5, 2.0
    _compare(loss, masks, classes, mask_labels, class_labels)


def test_batch_without_any_target_matches_reference():
    """No instance in the whole batch: the mask losses are exactly zero and only the "no object" class loss remains."""
    loss, masks, classes, mask_labels, class_labels = make_problem(7, L=2, n_tgt=(0, 0, 0))
    _compare(loss, masks, classes, mask_labels, class_labels)
    from weed_instance_segmentation_b200.criterion import convert_criterion
    mine = convert_criterion(copy.deepcopy(loss), sampler=grid_sample_sampler)
    out, gm, gc = run(mine, masks, classes, mask_labels, class_labels, seed=1)
    assert out["loss_mask"].item() == 0.0 and out["loss_dice"].item() == 0.0 and out["loss_cross_entropy"].item() > 0
    assert all(g is None or g.abs().max().item() == 0.0 for g in gm) and all(g.abs().max().item() > 0 for g in gc)


def test_mixed_size_targets_are_refused():
    """The reference pads targets to the largest size of the batch before sampling (M2F:612); the batched criterion
    samples planes in place, so it refuses such a batch instead of silently sampling other points."""
    from weed_instance_segmentation_b200.criterion import convert_criterion
    loss, masks, classes, mask_labels, class_labels = make_problem(3)
    mask_labels[1] = mask_labels[1][:, :-8, :-4].contiguous()
    mine = convert_criterion(copy.deepcopy(loss), sampler=grid_sample_sampler)
    with pytest.raises(ValueError, match="different sizes"):
        run(mine, masks, classes, mask_labels, class_labels, seed=1)
