"""`models.builder` of the reference (models/builder.py:13): EncoderDecoder."""
from sigma_b200.modules import EncoderDecoder  # noqa: F401
