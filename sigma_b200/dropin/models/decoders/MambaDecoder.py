"""`models.decoders.MambaDecoder` of the reference (models/decoders/MambaDecoder.py)."""
from sigma_b200.modules import FinalUpsample_X4, Mamba_up, MambaDecoder, PatchExpand, UpsampleExpand  # noqa: F401
