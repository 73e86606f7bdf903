"""`models.encoders.vmamba` of the reference (models/encoders/vmamba.py): the public names of the Sigma path."""
from sigma_b200.modules import (Backbone_VSSM, ChannelAttention, ChannelAttentionBlock, ConcatMambaFusionBlock,  # noqa: F401
                                ConMB_SS2D, Cross_Mamba_Attention_SSM, CrossMambaFusion_SS2D_SSM, CrossMambaFusionBlock,
                                CVSSDecoderBlock, Mlp, PatchMerging2D, Permute, SS2D, VSSBlock, VSSM)
from sigma_b200.ops import (CrossMerge, CrossMerge_multimodal, CrossScan, CrossScan_multimodal, SelectiveScan,  # noqa: F401
                            cross_selective_scan, cross_selective_scan_multimodal_k2)
