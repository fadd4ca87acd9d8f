"""Drop-in counterpart of the reference's ``nets`` package for the HNOSeg-XS hot path (CUDA only).

Same class names, constructor signatures, attribute names and ``state_dict`` keys as
``multimodal_3d_image_segmentation.nets`` (reference nets/__init__.py:11-12), so
``getattr(nets, model_name)(**model_args)`` in experiments/run.py:82-87 works unchanged.
"""
from . import custom_losses  # noqa: F401
from .architectures import HartleyMHABlock, HartleyMHASeg, NeuralOperatorSeg  # noqa: F401
from .dht import dht2, dht3, dhtn  # noqa: F401
from .fourier_operator import FourierOperator  # noqa: F401
from .hartley_mha import HartleyMultiHeadAttention  # noqa: F401
from .hartley_operator import HartleyOperator, get_reverse, hartley_conv  # noqa: F401
from .hnosegxs import HNOSegXS, HNOXSBlock, NeuralOperatorBlock, PadInverse, TransformCrop  # noqa: F401
from .nets_utils import ConvNormAct, init_weights_for_snn, spatial_padcrop  # noqa: F401
