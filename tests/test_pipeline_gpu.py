"""InferencePipeline (sigma_b200/pipeline.py): overlapped H2D / forward / D2H must return, for every submitted batch, exactly
what the direct module call returns for that batch (staging buffers, stream ordering, graph replay)."""
import contextlib
import io

import pytest
import torch

from helpers import cfg_tiny

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [True, False])
def test_pipeline_matches_direct_calls(use_graph):
    from sigma_b200 import modules as M
    from sigma_b200.pipeline import InferencePipeline
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)
    H, W, B = 64, 96, 2
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.EncoderDecoder(cfg_tiny(H, W), criterion=None).cuda().eval()
    pipe = InferencePipeline(model, B, H, W, use_graph=use_graph)
    nb = 5
    h_rgb = [torch.randn(B, 3, H, W).pin_memory() for _ in range(nb)]
    h_x = [torch.randn(B, 3, H, W).pin_memory() for _ in range(nb)]
    h_out = [torch.empty(pipe.out_shape).pin_memory() for _ in range(nb)]
    for i in range(nb):
        pipe.submit(h_rgb[i], h_x[i], h_out[i])
    pipe.drain()
    with torch.no_grad():
        for i in range(nb):
            ref = model(h_rgb[i].cuda(), h_x[i].cuda()).cpu()
            err = float((ref - h_out[i]).abs().max())
            # same kernels either way; cuDNN may pick another algorithm for the 4 dense 3x3 convs under graph capture
            assert err <= 1e-4 * max(1.0, float(ref.abs().max())), f"batch {i}: pipeline output differs from the direct call (max abs {err:.3e})"
            if i:   # and it is THIS batch's result, not a neighbour's
                assert float((h_out[i] - h_out[i - 1]).abs().max()) > 1e-3


def test_pipeline_requires_cuda_model():
    from sigma_b200 import modules as M
    from sigma_b200.pipeline import InferencePipeline
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.EncoderDecoder(cfg_tiny(64, 96), criterion=None).eval()
    with pytest.raises(RuntimeError):
        InferencePipeline(model, 1, 64, 96)


def test_pipeline_recaptures_after_weight_update():
    """The graph bakes in the pointers of the packed SSM tensors derived from the parameters: after an in-place weight change
    (load_state_dict) the pipeline must serve the NEW weights, not stale packed copies, and it refuses a model in train mode."""
    import procedural as P
    from sigma_b200 import modules as M
    from sigma_b200.pipeline import InferencePipeline
    torch.backends.cuda.matmul.allow_tf32 = True
    H, W, B = 64, 96, 1
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.EncoderDecoder(cfg_tiny(H, W), criterion=None).cuda().eval()
    P.fill_state_dict(model, 3)
    pipe = InferencePipeline(model, B, H, W)
    h_rgb, h_x = torch.randn(B, 3, H, W).pin_memory(), torch.randn(B, 3, H, W).pin_memory()
    out0, out1 = torch.empty(pipe.out_shape).pin_memory(), torch.empty(pipe.out_shape).pin_memory()
    pipe.submit(h_rgb, h_x, out0)
    pipe.drain()
    P.fill_state_dict(model, 4)                      # in-place update of every parameter (incl. x_proj_weight, A_logs, Ds, dt_projs_*)
    pipe.submit(h_rgb, h_x, out1)
    pipe.drain()
    with torch.no_grad():
        ref = model(h_rgb.cuda(), h_x.cuda()).cpu()
    assert float((out1 - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), "stale weights replayed after load_state_dict"
    assert float((out1 - out0).abs().max()) > 1e-3
    model.train()
    with pytest.raises(RuntimeError):
        InferencePipeline(model, B, H, W)
