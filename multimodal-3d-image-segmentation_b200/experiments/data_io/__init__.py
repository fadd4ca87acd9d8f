"""Device-side counterpart of the reference's experiments/data_io augmentation."""
from .dataset import ImageTransform, draw_transform  # noqa: F401
