"""Install the B200 operator into the reference's model.

The reference has no plugin mechanism; its model is HuggingFace's
``Mask2FormerForUniversalSegmentation`` (``/root/reference/models/model_utils.py:13-14``,
``models/mask2former/train.py:167-173``). Every pixel-decoder encoder layer calls the
module-level function ``multi_scale_deformable_attention`` (M2F:980), which Python resolves
as a module global at call time -- so rebinding that one name routes all six layers through
``libmsda_b200.so`` without touching weights or state-dict keys.

Usage, right after ``load_model`` / ``from_pretrained`` in the reference::

    import weed_instance_segmentation_b200 as wis
    wis.install()            # or: with wis.installed(): ...
"""
from __future__ import annotations

import contextlib

from .functional import multi_scale_deformable_attention as _b200_msda

_original = None


def _module():
    import transformers.models.mask2former.modeling_mask2former as m2f
    return m2f


def install() -> None:
    """Rebind ``modeling_mask2former.multi_scale_deformable_attention`` to the B200 operator."""
    global _original
    m2f = _module()
    if m2f.multi_scale_deformable_attention is _b200_msda:
        return
    _original = m2f.multi_scale_deformable_attention
    m2f.multi_scale_deformable_attention = _b200_msda


def uninstall() -> None:
    """Restore the reference implementation."""
    global _original
    if _original is not None:
        _module().multi_scale_deformable_attention = _original
        _original = None


def is_installed() -> bool:
    return _module().multi_scale_deformable_attention is _b200_msda


@contextlib.contextmanager
def installed():
    was = is_installed()
    install()
    try:
        yield
    finally:
        if not was:
            uninstall()
