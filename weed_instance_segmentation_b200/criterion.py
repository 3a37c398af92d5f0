"""Loss / matcher host path of the training step, batched over decoder layers.

The reference's ``Mask2FormerLoss.forward`` (M2F:739-783) runs once per decoder layer (10 times with auxiliary
losses); each run loops over the images of the batch in the Hungarian matcher (M2F:455-474: softmax, two
``sample_point`` calls, three matmuls and one ``.cpu()`` synchronisation per image) and then gathers the matched
prediction masks and a padded full-resolution copy of the target masks for the point-sampled mask losses
(M2F:689-737). At batch 8 that is 80 device synchronisations, ~2 500 kernel launches and ~3 GB of mask copies per
micro-batch -- 68 ms of a 183 ms forward on a B200 (profiles/r01_notes.md).

``B200Mask2FormerLoss`` computes the same losses with the same random numbers:

* every ``torch.rand`` of the reference is drawn first, in the reference's order and shapes, so that for a given
  generator state both implementations sample the same points;
* one ``point_sample`` launch (``csrc/point_sample.cu``) samples every (layer, image, query) prediction plane and
  every target plane at the matcher points, reading the planes where they are;
* the pairwise costs of all layers and images come from three batched matmuls; all cost matrices cross to the
  host in ONE copy and ``scipy.optimize.linear_sum_assignment`` runs on the slices;
* the uncertainty oversampling, the top-k and the final sampling run once for all layers (two more launches),
  with gradients flowing back into each layer's mask logits through ``point_sample``'s backward;
* the class loss of all layers is one weighted cross entropy.

It is a subclass that only overrides ``forward``: ``convert_criterion`` switches an existing ``Mask2FormerLoss``
instance over in place (same buffers, same ``state_dict`` keys, ``get_num_masks`` and the matcher weights reused).
"""
from __future__ import annotations

import functools

import numpy as np
import torch
import torch.nn.functional as F
from torch.profiler import record_function


def _host_int(values) -> np.ndarray:
    return np.asarray(values, dtype=np.int64).reshape(-1)


def _criterion_forward(self, masks_queries_logits, class_queries_logits, mask_labels, class_labels,
                       auxiliary_predictions=None):
    from scipy.optimize import linear_sum_assignment

    sampler = self._b200_sampler
    aux = list(auxiliary_predictions) if auxiliary_predictions is not None else []
    layer_masks = [masks_queries_logits] + [a["masks_queries_logits"] for a in aux]
    layer_classes = [class_queries_logits] + [a["class_queries_logits"] for a in aux]
    L = len(layer_masks)
    B, Q, h, w = layer_masks[0].shape
    device = layer_masks[0].device
    n_tgt = [int(c.shape[0]) for c in class_labels]
    n_match = [min(Q, n) for n in n_tgt]
    M, n_max = sum(n_match), max(max(n_tgt), 1)
    K_match, K = self.matcher.num_points, self.num_points
    k_over = int(K * self.oversample_ratio)
    k_unc = int(self.importance_sample_ratio * K)
    k_rand = K - k_unc

    with torch.autocast(device.type, enabled=False):
        # ---- 1. the reference's random draws, in its order (M2F:455 per image, then M2F:718 and :735 per layer)
        with torch.no_grad(), record_function("b200_loss::random_points"):
            pts_match, pts_over, pts_rand = [], [], []
            for _ in range(L):
                pts_match.extend(torch.rand(1, K_match, 2, device=device) for _ in range(B))
                pts_over.append(torch.rand(M, k_over, 2, device=device))
                if k_rand > 0:
                    pts_rand.append(torch.rand(M, k_rand, 2, device=device))

        preds = [m.reshape(B * Q, h, w) for m in layer_masks]                    # views: gradients reach the layers
        targets = [t if t.dtype in (torch.float32, torch.bfloat16, torch.uint8, torch.bool) else t.float()
                   for t in mask_labels]                                             # binary masks may stay 1 byte/pixel
        targets = [t.reshape(-1, *t.shape[-2:]) for t in targets]
        if len({tuple(t.shape[-2:]) for t in targets}) > 1:
            # the reference pads all targets to the largest size of the batch first (M2F:612) and samples the PADDED
            # extent; sampling every plane at its own extent would silently give other point labels
            raise ValueError("mask_labels of different sizes in one batch are not supported by the batched criterion "
                             "(the reference's collate stacks equally sized images, dataset_utils.py:45-53)")
        sources = preds + targets

        # ---- 2. matcher costs of every (layer, image) at once (M2F:440-470)
        with torch.no_grad(), record_function("b200_loss::matcher_costs"):
            lay, img = np.divmod(np.arange(L * B), B)
            p_src = np.repeat(lay, Q)
            p_plane = (np.repeat(img, Q) * Q + np.tile(np.arange(Q), L * B))
            p_crow = np.repeat(np.arange(L * B), Q)
            t_src = np.concatenate([np.full(n_tgt[i], L + i) for i in img]) if sum(n_tgt) else np.zeros(0, np.int64)
            t_plane = np.concatenate([np.arange(n_tgt[i]) for i in img]) if sum(n_tgt) else np.zeros(0, np.int64)
            t_crow = np.concatenate([np.full(n_tgt[i], li) for li, i in enumerate(img)]) if sum(n_tgt) else np.zeros(0, np.int64)
            sampled = sampler(sources, np.concatenate([p_src, t_src]), np.concatenate([p_plane, t_plane]),
                              torch.cat(pts_match), np.concatenate([p_crow, t_crow]))
            x = sampled[: L * B * Q].view(L * B, Q, K_match)
            t_pad = torch.zeros(L * B * n_max, K_match, device=device)
            if t_plane.size:
                dest = torch.from_numpy(t_crow * n_max + t_plane).to(device)
                t_pad[dest] = sampled[L * B * Q:]
            t_pad = t_pad.view(L * B, n_max, K_match)
            tt = t_pad.transpose(1, 2)
            # pair-wise sigmoid cross entropy (M2F:352-372) and dice (M2F:328-348)
            cost_mask = torch.bmm(F.softplus(-x) / K_match, tt) + torch.bmm(F.softplus(x) / K_match, 1 - tt)
            probs = x.sigmoid()
            cost_dice = 1 - (2 * torch.bmm(probs, tt) + 1) / (probs.sum(-1)[:, :, None] + t_pad.sum(-1)[:, None, :] + 1)
            cls_prob = torch.stack(layer_classes).float().softmax(-1).view(L * B, Q, -1)
            lab_pad = torch.zeros(B, n_max, dtype=torch.int64, device=device)
            for i, c in enumerate(class_labels):
                lab_pad[i, : n_tgt[i]] = c
            cost_class = -cls_prob.gather(2, lab_pad.repeat(L, 1)[:, None, :].expand(-1, Q, -1))
            cost = self.matcher.cost_mask * cost_mask + self.matcher.cost_class * cost_class \
                + self.matcher.cost_dice * cost_dice
            cost = torch.nan_to_num(cost.clamp(-1e10, 1e10), 0)
            cost_host = cost.cpu().numpy()                                      # the one synchronisation
        indices = []                                                            # [layer][image] -> (pred idx, target idx)
        with record_function("b200_loss::assignment"):
            for li in range(L * B):
                i = int(img[li])
                indices.append(linear_sum_assignment(cost_host[li][:, : n_tgt[i]]))

        # ---- 3. matched rows of every layer, in the reference's order (M2F:707-716)
        img_of_pair = np.concatenate([np.full(n_match[i], i) for i in range(B)]) if M else np.zeros(0, np.int64)
        pred_plane = np.zeros((L, M), np.int64)
        tgt_plane = np.zeros((L, M), np.int64)
        for l in range(L):
            if M:
                pred_plane[l] = np.concatenate([_host_int(indices[l * B + i][0]) + i * Q for i in range(B)])
                tgt_plane[l] = np.concatenate([_host_int(indices[l * B + i][1]) for i in range(B)])
        lay_of_row = np.repeat(np.arange(L), M)
        row_ids = np.arange(L * M)

        # ---- 4. importance sampling of the loss points (M2F:646-687), all layers in one pass
        with torch.no_grad(), record_function("b200_loss::importance_sampling"):
            coords_over = torch.cat(pts_over)                                                   # (L*M, k_over, 2)
            over = sampler(sources, lay_of_row, pred_plane.reshape(-1), coords_over, row_ids)     # (L*M, k_over)
            idx = torch.topk(-over.abs(), k=k_unc, dim=1)[1]
            coords = coords_over.gather(1, idx[:, :, None].expand(-1, -1, 2))
            if k_rand > 0:
                coords = torch.cat([coords, torch.cat(pts_rand)], dim=1)

        # ---- 5. point logits (with gradient) and point labels, mask + dice losses (M2F:727-737, :283-327)
        num_masks = self.get_num_masks(class_labels, device=class_labels[0].device)
        tgt_src = np.tile(L + img_of_pair, L)
        both = sampler(sources, np.concatenate([lay_of_row, tgt_src]),
                       np.concatenate([pred_plane.reshape(-1), tgt_plane.reshape(-1)]), coords,
                       np.concatenate([row_ids, row_ids]))
        point_logits, point_labels = both[: L * M], both[L * M:].detach()
        ce = F.binary_cross_entropy_with_logits(point_logits, point_labels, reduction="none")
        loss_mask = ce.mean(1).view(L, M).sum(1) / num_masks
        pr = point_logits.sigmoid()
        dice = 1 - (2 * (pr * point_labels).sum(-1) + 1) / (pr.sum(-1) + point_labels.sum(-1) + 1)
        loss_dice = dice.view(L, M).sum(1) / num_masks

        # ---- 6. class loss of every layer (M2F:571-604): weighted cross entropy, "no object" everywhere else
        logits = torch.stack(layer_classes).float()                                             # (L, B, Q, C+1)
        target_classes = torch.full((L * B * Q,), self.num_labels, dtype=torch.int64, device=device)
        if M:
            offsets = np.concatenate([[0], np.cumsum(n_tgt)[:-1]])
            dest = (lay_of_row * B * Q + pred_plane.reshape(-1))
            src = np.tile(offsets[img_of_pair], L) + tgt_plane.reshape(-1)
            target_classes[torch.from_numpy(dest).to(device)] = torch.cat(list(class_labels))[torch.from_numpy(src).to(device)]
        logp = logits.log_softmax(-1).view(L * B * Q, -1)
        nll = -logp.gather(1, target_classes[:, None]).squeeze(1)
        wt = self.empty_weight.float()[target_classes]
        loss_ce = (wt * nll).view(L, -1).sum(1) / wt.view(L, -1).sum(1)

    losses = {}
    for l in range(L):
        suffix = "" if l == 0 else f"_{l - 1}"
        # separate tensors: the caller scales them in place (M2F:2287-2290)
        losses[f"loss_mask{suffix}"] = loss_mask[l].clone()
        losses[f"loss_dice{suffix}"] = loss_dice[l].clone()
        losses[f"loss_cross_entropy{suffix}"] = loss_ce[l].clone()
    self.last_indices = indices
    return losses


@functools.lru_cache(maxsize=1)
def loss_class():
    """``B200Mask2FormerLoss``: a ``Mask2FormerLoss`` whose ``forward`` is the batched path above."""
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerLoss

    def default_sampler(*args):
        from .point_sample import point_sample
        return point_sample(*args)

    return type("B200Mask2FormerLoss", (Mask2FormerLoss,), {
        "forward": _criterion_forward, "_b200_sampler": staticmethod(default_sampler), "last_indices": None,
        "__doc__": __doc__, "__module__": __name__, "__qualname__": "B200Mask2FormerLoss",
    })


def __getattr__(name):
    # `criterion.B200Mask2FormerLoss` resolves lazily (transformers is imported on first use), which also lets
    # pickle / copy find the class of a converted module by name.
    if name == "B200Mask2FormerLoss":
        return loss_class()
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def convert_criterion(model_or_loss, sampler=None):
    """Switch a ``Mask2FormerLoss`` (or the ``.criterion`` of a model) to the batched path, in place.

    ``sampler`` replaces ``point_sample`` (same signature); tests use it to check the host logic on the CPU against
    the reference criterion. Returns the converted loss module.
    """
    loss = getattr(model_or_loss, "criterion", model_or_loss)
    loss.__class__ = loss_class()
    if sampler is not None:
        loss._b200_sampler = sampler
    return loss


def restore_criterion(model_or_loss):
    """Undo ``convert_criterion``."""
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerLoss
    loss = getattr(model_or_loss, "criterion", model_or_loss)
    loss.__class__ = Mask2FormerLoss
    loss.__dict__.pop("_b200_sampler", None)
    return loss
