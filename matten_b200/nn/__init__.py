from .conv import PointConv, PointConvWithActivation  # noqa: F401
from .embedding import EdgeLengthEmbedding, SpeciesEmbedding  # noqa: F401
from ._nequip import RadialBasisEdgeEncoding, SphericalHarmonicEdgeAttrs  # noqa: F401
from .nodewise import NodewiseLinear, NodewiseReduce, NodewiseSelect  # noqa: F401
from .sequential import Sequential  # noqa: F401
from .utils import ActivationLayer, NormalizationLayer, UVUTensorProduct  # noqa: F401
