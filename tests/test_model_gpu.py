"""The B200 operator dropped into the reference's model (HF Mask2Former) vs the unmodified model.

Same weights, same synthetic inputs; compares logits, loss, parameter gradients and the
post-processed per-instance masks (north_star: "per-instance mask IoU unchanged on the synthetic
eval set"). The reference path here is the stock HF module graph (M2F:798-837 grid_sample op).
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(decoder_layers=4, encoder_layers=6, num_queries=20, train_num_points=256)


@pytest.fixture(scope="module")
def models():
    pytest.importorskip("transformers")
    from weed_instance_segmentation_b200 import _cabi, build, train
    build.build()
    _cabi.load()
    ref = train.build_model("swin_tiny_test", num_labels=3, seed=0, **CFG).cuda()
    # A freshly initialised module has zero offset weights and integer-pixel biases (M2F:2116-2128), so every
    # sample sits exactly ON a grid line, where bilinear sampling is not differentiable and d/d(loc) is whatever
    # side floor() picks in the last bit. Move the biases off the grid (as any trained checkpoint is) so that the
    # location gradients are well defined and comparable between implementations.
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if name.endswith("self_attn.sampling_offsets.bias"):
                p.add_(0.1 + 0.3 * torch.rand(p.shape, generator=g, device="cuda"))
    fn = copy.deepcopy(ref)
    mod = copy.deepcopy(ref)
    full = copy.deepcopy(ref)
    return ref, fn, mod, full


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def _forward(model, batch, patched):
    import weed_instance_segmentation_b200 as wis
    torch.manual_seed(1234)  # the loss samples random points (M2F:619-631)
    if patched:
        with wis.installed():
            return model(pixel_values=batch["pixel_values"], mask_labels=batch["mask_labels"],
                         class_labels=batch["class_labels"])
    assert not wis.is_installed()
    return model(pixel_values=batch["pixel_values"], mask_labels=batch["mask_labels"], class_labels=batch["class_labels"])


@pytest.mark.parametrize("size", [(128, 160), (97, 130)])
def test_logits_loss_and_grads_match_reference(models, size):
    from weed_instance_segmentation_b200 import _cabi, modules, synth
    from weed_instance_segmentation_b200.criterion import convert_criterion
    ref, fn, mod, full = models
    modules.convert_pixel_decoder(mod)
    modules.convert_pixel_decoder(full)
    convert_criterion(full)  # + batched loss / matcher: same random points, so the loss is comparable to 1e-4 as well
    batch = synth.collate_batch(2, size[0], size[1], num_classes=3, max_instances=4, seed=3, device="cuda")
    outs = {}
    for name, model, patched in (("ref", ref, False), ("fn", fn, True), ("mod", mod, True), ("full", full, True)):
        model.train()
        model.zero_grad(set_to_none=True)
        _cabi.launch_count(reset=True)
        out = _forward(model, batch, patched)
        out.loss.backward()
        outs[name] = (out, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
                      _cabi.launch_count())
    assert outs["ref"][2] == 0 and outs["fn"][2] >= 6 * 2 and outs["mod"][2] >= 6 * 2  # 6 layers fwd + bwd kernels
    assert outs["full"][2] >= 6 * 2 + 4  # + three sampling launches and their backward
    for name in ("fn", "mod", "full"):
        o, g, _ = outs[name]
        r, rg, _ = outs["ref"]
        # the op agrees with the reference to ~1e-6 per call (test_msda_gpu); six encoder layers, LayerNorms and
        # four decoder layers amplify that to ~1.5e-4 at the logits (measured), so the model-level bar is 5e-4
        assert _rel(o.masks_queries_logits, r.masks_queries_logits) < 5e-4, name
        assert _rel(o.class_queries_logits, r.class_queries_logits) < 5e-4, name
        assert abs(o.loss.item() - r.loss.item()) < 1e-4 * abs(r.loss.item()), name
        assert set(g) == set(rg)
        worst = max((_rel(g[k], rg[k]), k) for k in rg if rg[k].abs().max() > 1e-6)
        # measured worst case 2.7e-3 (layers.3.fc1.weight): fp32 round-off of the logits amplified by the loss
        assert worst[0] < 1e-2, (name, worst)


def test_instance_masks_unchanged(models):
    """post_process_instance_segmentation (the reference's eval path, models/metrics.py:58-63) yields the
    same instances: per-instance mask IoU == 1 between the stock and the B200 model."""
    from transformers import Mask2FormerImageProcessor
    import weed_instance_segmentation_b200 as wis
    from weed_instance_segmentation_b200 import synth
    ref, fn, _, _ = models
    proc = Mask2FormerImageProcessor()
    ious, differing, explained = [], 0, 0
    for seed in range(3):
        batch = synth.collate_batch(2, 128, 160, num_classes=3, max_instances=4, seed=20 + seed, device="cuda")
        with torch.no_grad():
            ref.eval(), fn.eval()
            o_ref = ref(pixel_values=batch["pixel_values"])
            with wis.installed():
                o_b200 = fn(pixel_values=batch["pixel_values"])
        kw = dict(threshold=0.0, mask_threshold=0.5, target_sizes=batch["target_sizes"])
        p_ref = proc.post_process_instance_segmentation(o_ref, **kw)
        p_new = proc.post_process_instance_segmentation(o_b200, **kw)
        # pixels where ANY query's mask logit sits within 1e-4 of the largest logit magnitude from the 0.5-probability
        # threshold: only those may change owner.  The post-processing (M2FIP:605-725) resamples the logits bilinearly to
        # 384 x 384, thresholds there, and carries the binary masks to the target size with nearest-neighbour
        # interpolation; the ambiguity mask takes the same route.
        logits = o_ref.masks_queries_logits
        tol = 1e-4 * logits.abs().max().item()
        for i, (a, b) in enumerate(zip(p_ref, p_new)):
            sa, sb = a["segmentation"], b["segmentation"]
            up = torch.nn.functional.interpolate(logits[i][None], size=(384, 384), mode="bilinear", align_corners=False)[0]
            amb = (up.abs() < tol).any(0).float()[None, None]
            ambiguous = (torch.nn.functional.interpolate(amb, size=tuple(sa.shape), mode="nearest")[0, 0] > 0).to(sa.device)
            diff = sa != sb
            differing += int(diff.sum())
            explained += int((diff & ambiguous).sum())
            ids = [int(k) for k in sa.unique().tolist() if k >= 0]
            assert len(a["segments_info"]) == len(b["segments_info"])
            for k in ids:
                ma, mb = (sa == k) & ~ambiguous, (sb == k) & ~ambiguous
                ious.append((ma & mb).sum().item() / max((ma | mb).sum().item(), 1))
    assert ious, "no instances produced; lower the threshold"
    # per-instance mask IoU is exactly 1 outside the threshold-ambiguous pixels, and every differing pixel is one of them
    assert differing == explained, (differing, explained)
    assert min(ious) == 1.0, (min(ious), sum(ious) / len(ious))


def test_config1_geometry_swin_t_inference_matches(models):
    """BASELINE.json configs[0] geometry: the full-size Swin-T model, batch 1, 512x512, fp32 -- stock model vs the B200
    operator installed (M2F:980 rebound), same weights."""
    import weed_instance_segmentation_b200 as wis
    from weed_instance_segmentation_b200 import train
    model = train.build_model("swin_t", num_labels=3, seed=0).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(0)
    pixel_values = torch.randn(1, 3, 512, 512, generator=g, device="cuda")
    with torch.no_grad():
        want = model(pixel_values=pixel_values)
        with wis.installed():
            got = model(pixel_values=pixel_values)
    # six encoder layers + ten decoder layers of a random-init model amplify the op's ~1e-6 fp32 round-off (and cuDNN's
    # TF32 convolutions see slightly different inputs): measured 1.8e-3 on the mask logits (max-norm relative)
    assert _rel(got.masks_queries_logits, want.masks_queries_logits) < 5e-3
    assert _rel(got.class_queries_logits, want.class_queries_logits) < 5e-3


def test_bf16_autocast_forward_close(models):
    from weed_instance_segmentation_b200 import synth
    import weed_instance_segmentation_b200 as wis
    ref, fn, _, _ = models
    batch = synth.collate_batch(2, 128, 160, num_classes=3, max_instances=4, seed=5, device="cuda")
    ref.eval(), fn.eval()
    with torch.no_grad():
        o_f32 = ref(pixel_values=batch["pixel_values"])
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o_ref = ref(pixel_values=batch["pixel_values"])
            with wis.installed():
                o_new = fn(pixel_values=batch["pixel_values"])
    # every layer of both autocast runs carries bf16 rounding; the B200 path must sit as close to the
    # fp32 model as the stock autocast path does (the op alone is held to 2e-2 in test_msda_gpu)
    for key in ("masks_queries_logits", "class_queries_logits"):
        e_ref = _rel(getattr(o_ref, key), getattr(o_f32, key))
        e_new = _rel(getattr(o_new, key), getattr(o_f32, key))
        assert e_new <= 2.0 * e_ref + 1e-2, (key, e_new, e_ref)
