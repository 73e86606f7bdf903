def _missing(*a, **k):
    raise ImportError("fvcore is not installed: FLOP counting (models/*.flops) is unavailable; training / evaluation do not need it")


FlopCountAnalysis = flop_count_str = flop_count = parameter_count = _missing
