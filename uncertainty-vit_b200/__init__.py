"""b200-uncertainty-vit: the ViT-B/16 / ViT-L/16 data2vec training + uncertainty-inference hot path of
fx-erick/uncertainty-vit as hand-written sm_100a CUDA behind the reference's model-registry boundary.

Import as `uncertainty_vit_b200` (the root-level loader maps the hyphenated directory to that module name)."""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
