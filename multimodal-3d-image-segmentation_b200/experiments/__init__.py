"""Device-side counterparts of the array helpers in the reference's experiments/utils.py and of the augmentation in
experiments/data_io/dataset.py."""
from . import data_io  # noqa: F401
from .utils import normalize_modalities, to_categorical  # noqa: F401
