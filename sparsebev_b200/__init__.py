"""sparsebev_b200: B200-native (sm_100a) implementation of SparseBEV's decoder hot path --
adaptive spatio-temporal sampling (msmv_sampling) + scale-adaptive self-attention + adaptive mixing --
behind the reference's own plugin surface.  See DESIGN.md / INTEGRATION.md."""
from .wrapper import MSMV_CUDA, msmv_sampling, msmv_sampling_pytorch, MSMVSamplingC2345, MSMVSamplingC23456  # noqa: F401
from .sampling import sampling_4d, make_sample_points  # noqa: F401
from .transformer import (SparseBEVTransformer, SparseBEVTransformerDecoder, SparseBEVTransformerDecoderLayer,  # noqa: F401
                          SparseBEVSelfAttention, SparseBEVSampling, AdaptiveMixing)
from .head import SparseBEVHead  # noqa: F401
from .coder import NMSFreeCoder, denormalize_bbox  # noqa: F401

__version__ = '0.1.0'
