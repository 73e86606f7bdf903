"""Stand-in for timm==0.4.12 (requirements.txt:10): the three names the reference imports from timm.models.layers."""
