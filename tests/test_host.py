"""Host-side logic that needs no GPU: synthetic generator, query order, level starts."""
import numpy as np
import pytest
import torch

from weed_instance_segmentation_b200 import functional as F
from weed_instance_segmentation_b200 import synth


def test_pixel_decoder_shapes_match_survey():
    assert synth.pixel_decoder_shapes(1024, 1024) == [(32, 32), (64, 64), (128, 128)]
    assert synth.pixel_decoder_shapes(512, 512) == [(16, 16), (32, 32), (64, 64)]
    assert synth.pixel_decoder_shapes(966, 1296) == [(31, 41), (61, 81), (121, 162)]
    assert synth.pixel_decoder_shapes(2048, 2048) == [(64, 64), (128, 128), (256, 256)]


@pytest.mark.parametrize("shapes", [[(32, 32), (64, 64), (128, 128)], [(31, 41), (61, 81), (121, 162)], [(1, 1), (1, 7), (5, 1)]])
@pytest.mark.parametrize("tile", [4, 8])
def test_query_order_is_a_levelwise_permutation(shapes, tile):
    order = F.query_order_2d(shapes, tile, "cpu").numpy()
    S = sum(h * w for h, w in shapes)
    assert order.dtype == np.int32 and order.shape == (S,)
    assert np.array_equal(np.sort(order), np.arange(S))
    start = 0
    for h, w in shapes:  # each level's block of the order stays inside the level
        blk = order[start:start + h * w]
        assert blk.min() == start and blk.max() == start + h * w - 1
        start += h * w
    # first tile of the last level is a compact patch
    h, w = shapes[-1]
    first = order[S - h * w: S - h * w + min(tile, w) * min(tile, h)] - (S - h * w)
    assert (first % w).max() < tile and (first // w).max() < tile


def test_level_start_derivation_and_override():
    shapes = [(2, 3), (4, 5)]
    assert F._level_start(shapes, None) == [0, 6]
    assert F._level_start(shapes, torch.tensor([0, 6])) == [0, 6]
    assert F._level_start(shapes, [0, 10]) == [0, 10]


def test_reference_points_match_hf():
    pytest.importorskip("transformers")
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerPixelDecoderEncoderOnly as Enc
    shapes = [(3, 4), (6, 7)]
    vr = torch.ones(1, len(shapes), 2)
    ref = Enc.get_reference_points(shapes, vr, "cpu")[0]
    assert torch.allclose(ref, synth.reference_points(shapes), atol=1e-7)


def test_init_offsets_match_hf_init():
    pytest.importorskip("transformers")
    from transformers import Mask2FormerConfig
    from transformers.models.mask2former.modeling_mask2former import (
        Mask2FormerPixelDecoderEncoderMultiscaleDeformableAttention as Attn, Mask2FormerPreTrainedModel)
    m = Attn(256, 8, 3, 4)
    Mask2FormerPreTrainedModel._init_weights(type("X", (), {"config": Mask2FormerConfig()})(), m)
    assert torch.allclose(m.sampling_offsets.bias.view(8, 3, 4, 2), synth.init_offsets(8, 3, 4), atol=1e-6)


@pytest.mark.parametrize("dist", ["init", "trained", "adversarial"])
def test_msda_inputs_contract(dist):
    x = synth.msda_inputs(2, [(2, 3), (4, 5)], dist=dist, seed=3, value_dtype=torch.bfloat16)
    assert x["value"].shape == (2, 26, 8, 32) and x["value"].dtype == torch.bfloat16
    assert x["sampling_locations"].shape == (2, 26, 8, 2, 4, 2) and x["sampling_locations"].dtype == torch.float32
    assert x["attention_weights"].shape == (2, 26, 8, 2, 4)
    assert torch.allclose(x["attention_weights"].float().sum((-1, -2)), torch.ones(2, 26, 8), atol=2e-2)
    assert x["level_start_index"].tolist() == [0, 6]
    y = synth.msda_inputs(2, [(2, 3), (4, 5)], dist=dist, seed=3, value_dtype=torch.bfloat16)
    assert torch.equal(x["sampling_locations"], y["sampling_locations"])  # seeded


def test_collate_batch_matches_reference_collate_fn_schema():
    b = synth.collate_batch(2, 64, 96, num_classes=3, max_instances=5, seed=1)
    # keys of /root/reference/datasets/dataset_utils.py:45-53
    assert list(b) == ["pixel_values", "mask_labels", "class_labels", "target_sizes", "original_maps", "id_mappings", "file_names"]
    assert b["pixel_values"].shape == (2, 3, 64, 96) and b["pixel_values"].dtype == torch.float32
    for m, c, o, idm in zip(b["mask_labels"], b["class_labels"], b["original_maps"], b["id_mappings"]):
        assert m.dtype == torch.float32 and m.shape[1:] == (64, 96) and m.shape[0] == c.shape[0] >= 1
        assert c.dtype == torch.int64 and int(c.max()) < 3
        assert o.dtype == torch.int32 and o.shape == (64, 96)
        assert len(idm) == m.shape[0]
        # /root/reference/datasets/pheno_bench/dataset.py:85: 255 where there is no instance, ids from 1 elsewhere
        assert set(o.unique().tolist()) == set(idm) | {255}
        for j, k in enumerate(sorted(idm)):
            assert torch.equal(m[j] > 0, o == k)
        assert set(m.unique().tolist()) <= {0.0, 1.0}
