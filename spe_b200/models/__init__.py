"""Mirror of the reference's `models` package (models/__init__.py:1-14): build_model(args)."""
from .conditional_detr import build


def build_model(args):
    return build(args)
