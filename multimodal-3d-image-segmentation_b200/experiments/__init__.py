"""Device-side counterparts of the array helpers in the reference's experiments/utils.py."""
from .utils import normalize_modalities, to_categorical  # noqa: F401
