"""`models.encoders.dual_vmamba` of the reference (models/encoders/dual_vmamba.py)."""
from sigma_b200.modules import RGBXTransformer, vssm_base, vssm_small, vssm_tiny  # noqa: F401
