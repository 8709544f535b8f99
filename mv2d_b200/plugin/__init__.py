"""Registry surface of the reference plugin, re-hosted on libmv2d_b200.

Importing this package registers, under the reference's own type names
(mmdet3d_plugin/__init__.py:10-19), modules with the reference's
constructor arguments, parameter names and shapes -- so ``configs/mv2d/*`` build unchanged and a
reference ``state_dict`` loads with ``strict=True`` -- whose forward runs the sm_100a kernels
through the C ABI.  The modules only HOLD parameters; arithmetic happens in
``mv2d_b200/csrc``.
"""
from .modules import (FPN, HungarianAssigner3D, PE, BoxCorrelation, CrossAttentionBoxHead, FlattenMHSelfAttention,  # noqa: F401
                      MV2D, MV2DT, MV2DHead, MV2DSHead, MV2DTHead, MV2DTransformer, NMSFreeCoder,
                      PETRMultiheadAttention, PETRTransformerDecoder, PETRTransformerDecoderLayer,
                      QueryGenerator, SinePositionalEncoding3D, SingleRoIExtractor)
