"""Synthetic inputs shaped like the reference's data, for tests and benchmarks.

Two generators:

* :func:`msda_inputs` -- op-level tensors for the MSDeformAttn hot path, with the
  three sampling-location distributions of SURVEY.md section 8(d):
  ``init`` (what a freshly initialised module produces, M2F:2116-2135: offsets are
  the 8 compass directions x (p+1) px, plus jitter), ``trained`` (offsets
  ~ N(0, 4 px)) and ``adversarial`` (locations uniform in [-0.1, 1.1], no locality).
* :func:`collate_batch` -- a training batch with exactly the keys and dtypes that
  ``/root/reference/datasets/dataset_utils.py:32-53`` (``collate_fn``) builds from the
  per-sample dicts of e.g. ``/root/reference/datasets/pheno_bench/dataset.py:127-135``.

There is no network and there are no datasets in this build; everything is seeded
random data of the right shape.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch

# BASELINE.json configs -> (input H, input W); the pixel decoder sees strides 32/16/8,
# coarsest level first (M2F:1303 iterates features[::-1][:3]).
CONFIG_IMAGE_SIZES = {
    "C1": (512, 512),
    "C2": (1024, 1024),
    "C3": (966, 1296),
    "C4": (1024, 1024),
    "C5": (2048, 2048),
}


def pixel_decoder_shapes(height: int, width: int) -> list[tuple[int, int]]:
    """Spatial shapes the Mask2Former pixel decoder feeds to MSDeformAttn.

    Swin patch-embed (stride 4, pads to a multiple of 4) followed by three 2x patch
    mergings that pad odd sizes up; the decoder uses strides 32, 16, 8 (coarsest first).
    A raw 966x1296 image gives (31,41) (61,81) (121,162), as probed in SURVEY.md section 8.
    """
    h, w = math.ceil(height / 4), math.ceil(width / 4)
    out = []
    for _ in range(3):
        h, w = (h + 1) // 2, (w + 1) // 2
        out.append((h, w))
    return out[::-1]


def level_start_index(shapes: Sequence[tuple[int, int]]) -> list[int]:
    """M2F:1321: exclusive prefix sum of H_l*W_l."""
    out, acc = [], 0
    for h, w in shapes:
        out.append(acc)
        acc += int(h) * int(w)
    return out


def reference_points(shapes: Sequence[tuple[int, int]], device="cpu", dtype=torch.float32) -> torch.Tensor:
    """M2F:1095-1125 with valid_ratios == 1: ``(S, L, 2)``, last dim (x, y).

    Every query's reference point is the centre of its own pixel, replicated for all levels.
    """
    refs = []
    for h, w in shapes:
        ry, rx = torch.meshgrid(
            torch.linspace(0.5, h - 0.5, h, dtype=dtype, device=device),
            torch.linspace(0.5, w - 0.5, w, dtype=dtype, device=device),
            indexing="ij",
        )
        refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
    ref = torch.cat(refs, 0)
    return ref[:, None, :].expand(-1, len(shapes), -1).contiguous()


def init_offsets(num_heads: int, num_levels: int, num_points: int) -> torch.Tensor:
    """The ``sampling_offsets.bias`` a fresh module gets (M2F:2118-2128): ``(H, L, P, 2)`` in pixels."""
    thetas = torch.arange(num_heads, dtype=torch.int64).float() * (2.0 * math.pi / num_heads)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2)
    grid = grid.repeat(1, num_levels, num_points, 1)
    for i in range(num_points):
        grid[:, :, i, :] *= i + 1
    return grid


def msda_inputs(
    batch: int,
    shapes: Sequence[tuple[int, int]],
    num_heads: int = 8,
    head_dim: int = 32,
    num_points: int = 4,
    dist: str = "init",
    seed: int = 0,
    device="cpu",
    value_dtype=torch.float32,
    attn_dtype=None,
    num_queries: int | None = None,
    with_grad_out: bool = True,
) -> dict:
    """Op-level inputs: value (B,S,H,D), loc (B,Q,H,L,P,2) fp32, attn (B,Q,H,L,P), grad_out (B,Q,H*D)."""
    shapes = [(int(h), int(w)) for h, w in shapes]
    L, P, H, D = len(shapes), num_points, num_heads, head_dim
    S = sum(h * w for h, w in shapes)
    Q = S if num_queries is None else int(num_queries)
    attn_dtype = value_dtype if attn_dtype is None else attn_dtype
    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def randn(*s):
        return torch.randn(*s, generator=g, device=device, dtype=torch.float32)

    value = randn(batch, S, H, D).to(value_dtype)
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=device)  # (L,2) = (W,H)
    if dist == "adversarial":
        loc = torch.rand(batch, Q, H, L, P, 2, generator=g, device=device) * 1.2 - 0.1
    else:
        if Q == S:
            ref = reference_points(shapes, device=device)  # (S,L,2)
        else:
            ref = torch.rand(Q, 1, 2, generator=g, device=device).expand(-1, L, -1)
        if dist == "init":
            off = init_offsets(H, L, P).to(device)[None, None] + 0.5 * randn(batch, Q, H, L, P, 2)
        elif dist == "trained":
            off = 4.0 * randn(batch, Q, H, L, P, 2)
        else:
            raise ValueError(f"unknown dist {dist!r}")
        # M2F:963-971: loc = ref[:, :, None, :, None, :] + off / (W_l, H_l)
        loc = ref[None, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    attn = torch.softmax(randn(batch, Q, H, L * P), -1).view(batch, Q, H, L, P).to(attn_dtype)
    out = {
        "value": value,
        "spatial_shapes": shapes,
        "level_start_index": torch.tensor(level_start_index(shapes), dtype=torch.int64, device=device),
        "sampling_locations": loc.contiguous(),
        "attention_weights": attn.contiguous(),
    }
    if with_grad_out:
        out["grad_out"] = randn(batch, Q, H * D).to(value_dtype)
    return out


def collate_batch(
    batch: int,
    height: int,
    width: int,
    num_classes: int = 3,
    max_instances: int = 20,
    seed: int = 0,
    device="cpu",
    mask_dtype: torch.dtype = torch.float32,
    pin_memory: bool = False,
) -> dict:
    """A batch in the reference's ``collate_fn`` layout (dataset_utils.py:45-53).

    ``mask_dtype=torch.uint8`` keeps the binary masks at one byte per pixel (a quarter of the host->device bytes;
    the batched criterion samples them as they are, the stock criterion needs float32); ``pin_memory`` page-locks
    ``pixel_values`` and the masks so that ``Trainer.prefetch`` copies them asynchronously.

    ``pixel_values`` (B,3,H,W) float32 stacked; ``mask_labels`` list of (N_i,H,W) float32 binary
    masks; ``class_labels`` list of (N_i,) int64; ``target_sizes`` list of (h,w);
    ``original_maps`` list of (H,W) int32 instance maps, instance ids from 1, 255 = background / ignore as the
    reference builds them (``/root/reference/datasets/pheno_bench/dataset.py:85``);
    ``id_mappings`` list of {instance_id: class_id}; ``file_names`` list of str.
    Instances are random axis-aligned ellipses ("blobs"), N_i ~ U{1..max_instances}.
    """
    if not 1 <= max_instances < 255:
        raise ValueError("max_instances must lie in [1, 254]: 255 marks 'no instance' in original_maps")
    g = torch.Generator()
    g.manual_seed(seed)
    pixel_values = torch.randn(batch, 3, height, width, generator=g)
    yy = torch.arange(height, dtype=torch.float32)[:, None]
    xx = torch.arange(width, dtype=torch.float32)[None, :]
    mask_labels, class_labels, maps, mappings, names, sizes = [], [], [], [], [], []
    for i in range(batch):
        n = int(torch.randint(1, max_instances + 1, (1,), generator=g))
        inst = torch.full((height, width), 255, dtype=torch.int32)  # 255 = no instance (reference: np.full(..., 255))
        mapping = {}
        for k in range(1, n + 1):
            cy = float(torch.rand(1, generator=g)) * height
            cx = float(torch.rand(1, generator=g)) * width
            ry = (0.03 + 0.12 * float(torch.rand(1, generator=g))) * height
            rx = (0.03 + 0.12 * float(torch.rand(1, generator=g))) * width
            blob = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            inst[blob] = k
            mapping[k] = int(torch.randint(0, num_classes, (1,), generator=g))
        ids = [k for k in range(1, n + 1) if bool((inst == k).any())]
        if not ids:  # every blob overwritten or off-image: keep one pixel so N_i >= 1
            inst[height // 2, width // 2] = 1
            mapping.setdefault(1, 0)
            ids = [1]
        masks = torch.stack([(inst == k) for k in ids]).to(mask_dtype)
        mask_labels.append(masks.pin_memory() if pin_memory else masks.to(device))
        class_labels.append(torch.tensor([mapping[k] for k in ids], dtype=torch.int64, device=device))
        maps.append(inst)
        mappings.append({k: mapping[k] for k in ids})
        names.append(f"synthetic_{seed}_{i}.png")
        sizes.append((height, width))
    return {
        "pixel_values": pixel_values.pin_memory() if pin_memory else pixel_values.to(device),
        "mask_labels": mask_labels,
        "class_labels": class_labels,
        "target_sizes": sizes,
        "original_maps": maps,
        "id_mappings": mappings,
        "file_names": names,
    }
