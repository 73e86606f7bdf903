"""Checkpoint compatibility: the reference's `load_model` (utils/pyt_utils.py:155-192) for whole
EncoderDecoder checkpoints, next to the VMamba-pretraining renames that `sigma_b200.modules.VSSM._load_from_state_dict`
applies for backbone checkpoints (vmamba.py:2111-2147: patch_embed.proj/norm -> patch_embed.0/2, ln_1 -> norm,
self_attention -> op, norm/head -> classifier.*) and `Backbone_VSSM.load_pretrained` (vmamba.py:2181-2191)."""
import time
from collections import OrderedDict

import torch


def load_model(model, model_file, is_restore=False, logger=None):
    """utils/pyt_utils.py:155-192: `model_file` is a path or a state_dict; a dict with a 'model' / 'state_dict' / 'module'
    entry is unwrapped; `is_restore` re-adds the DDP 'module.' prefix; loading is strict, as in the reference."""
    t_start = time.time()
    if model_file is None:
        return model
    if isinstance(model_file, str):
        state_dict = torch.load(model_file, map_location="cpu")
        if "model" in state_dict.keys():
            state_dict = state_dict["model"]
        elif "state_dict" in state_dict.keys():
            state_dict = state_dict["state_dict"]
        elif "module" in state_dict.keys():
            state_dict = state_dict["module"]
    else:
        state_dict = model_file
    t_ioend = time.time()
    if is_restore:
        state_dict = OrderedDict(("module." + k, v) for k, v in state_dict.items())
    model.load_state_dict(state_dict, strict=True)
    # the fused inference path caches packed SSM tensors keyed by parameter versions: load_state_dict copies in place and bumps them
    if logger is not None:
        logger.info("Load model, Time usage:\n\tIO: {}, initialize parameters: {}".format(t_ioend - t_start, time.time() - t_ioend))
    return model
