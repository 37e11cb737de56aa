from .backbone import D2SwinTransformer, SwinTransformer  # noqa: F401
from .pixel_decoder import MSDeformAttnPixelDecoder  # noqa: F401
from .decoder import VideoMultiScaleMaskedTransformerDecoderUniVS  # noqa: F401
from .head import MaskFormerHead  # noqa: F401
