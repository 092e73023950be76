from ._contract import B200Contracter, Contracter  # noqa: F401
from ._mlp import ScalarMLPFunction  # noqa: F401
from ._modules import (  # noqa: F401
    Allegro_Module,
    EdgeLengthNormalizer,
    EdgewiseReduce,
    MakeWeightedChannels,
    PerTypeScaleShift,
    ProductTypeEmbedding,
    TwoBodyBesselScalarEmbed,
    TwoBodySplineScalarEmbed,
    PerClassSpline,
    TwoBodySphericalHarmonicTensorEmbed,
)
