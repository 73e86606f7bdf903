"""ORACLE — test infrastructure only.  CPU (torch fp32) restatement of the reference's model-forward
hot path, written functionally over a reference-format state_dict.  Never imported by sigma_b200/.

Each function cites the reference lines it follows (paths relative to zifuwan/Sigma @ 5c619c6).
Pinned by tests/test_oracle.py against tests/golden/*.npz, which were produced by running the
UNMODIFIED reference modules (tests/golden/make_golden.py); the selective scan itself goes through
the C oracle (oracle/selective_scan_ref.c).
"""
import math
import re

import numpy as np
import torch
import torch.nn.functional as F

from . import scan_oracle

EPS = 1e-5  # every LayerNorm on the Sigma path ends up with the nn.LayerNorm default


def selective_scan(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False):
    """selective_scan_interface.py:86-131 (selective_scan_ref), via the C oracle."""
    out = scan_oracle.scan_fwd(u.numpy(), delta.numpy(), A.numpy(), B.numpy(), C.numpy(),
                               None if D is None else D.numpy(),
                               None if delta_bias is None else delta_bias.numpy(), delta_softplus,
                               nthreads=torch.get_num_threads())
    return torch.from_numpy(out)


def ln(x, sd, pre):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + ".weight"], sd[pre + ".bias"], EPS)


def cross_scan(x):
    """vmamba.py:80-89.  (B,C,H,W) -> (B,4,C,L)"""
    B, C, H, W = x.shape
    xs = x.new_empty((B, 4, C, H * W))
    xs[:, 0] = x.flatten(2, 3)
    xs[:, 1] = x.transpose(2, 3).flatten(2, 3)
    xs[:, 2:4] = torch.flip(xs[:, 0:2], dims=[-1])
    return xs


def cross_merge(ys, H, W):
    """vmamba.py:100-108.  (B,4,D,L) -> (B,D,L)"""
    B, K, D, L = ys.shape
    ys = ys[:, 0:2] + ys[:, 2:4].flip(dims=[-1])
    return ys[:, 0] + ys[:, 1].view(B, D, W, H).transpose(2, 3).contiguous().view(B, D, L)


def _core(xs, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds):
    """Shared middle of cross_selective_scan (vmamba.py:195-215) and its multimodal twin (:401-421)."""
    B, K, D, L = xs.shape
    R = dt_projs_weight.shape[2]
    N = A_logs.shape[1]
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, x_proj_weight)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = torch.einsum("bkrl,kdr->bkdl", dts, dt_projs_weight)
    As = -torch.exp(A_logs.float())
    ys = selective_scan(xs.reshape(B, K * D, L), dts.reshape(B, K * D, L).contiguous(), As, Bs.contiguous(),
                        Cs.contiguous(), Ds.float(), dt_projs_bias.reshape(-1).float(), True)
    return ys.view(B, K, D, L)


def ss2d(x, sd, pre):
    """SS2D.forward (vmamba.py:1067-1089) with forward_corev2 = cross_selective_scan (vmamba.py:165-226)."""
    B, H, W, _ = x.shape
    D = sd[pre + ".conv2d.weight"].shape[0]
    xz = F.linear(x, sd[pre + ".in_proj.weight"])
    xi, z = xz.chunk(2, dim=-1)
    xi = xi.permute(0, 3, 1, 2).contiguous()
    xi = F.silu(F.conv2d(xi, sd[pre + ".conv2d.weight"], sd[pre + ".conv2d.bias"], padding=1, groups=D))
    ys = _core(cross_scan(xi), sd[pre + ".x_proj_weight"], sd[pre + ".dt_projs_weight"],
               sd[pre + ".dt_projs_bias"], sd[pre + ".A_logs"], sd[pre + ".Ds"])
    y = cross_merge(ys, H, W).transpose(1, 2).contiguous().view(B, H, W, D)
    y = ln(y, sd, pre + ".out_norm")
    y = y * F.silu(z)
    return F.linear(y, sd[pre + ".out_proj.weight"])


def vss_block(x, sd, pre):
    """VSSBlock._forward (vmamba.py:1712-1716) with mlp_ratio = 0 (dual_vmamba.py:119)."""
    return x + ss2d(ln(x, sd, pre + ".norm"), sd, pre + ".op")


def patch_merging(x, sd, pre):
    """PatchMerging2D (vmamba.py:619-636)."""
    H, W = x.shape[1:3]
    if (W % 2) or (H % 2):
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    return F.linear(ln(x, sd, pre + ".norm"), sd[pre + ".reduction.weight"])


def _count(sd, pattern):
    idx = set()
    for k in sd:
        m = re.match(pattern, k)
        if m:
            idx.add(int(m.group(1)))
    return len(idx)


def backbone(x, sd, pre):
    """Backbone_VSSM.forward (vmamba.py:2193-2212); patch_embed vmamba.py:1965-1969."""
    x = F.conv2d(x, sd[pre + ".patch_embed.0.weight"], sd[pre + ".patch_embed.0.bias"], stride=4)
    x = ln(x.permute(0, 2, 3, 1), sd, pre + ".patch_embed.2")
    outs = []
    for i in range(4):
        nblk = _count(sd, re.escape(f"{pre}.layers.{i}.blocks.") + r"(\d+)\.norm\.weight")
        for j in range(nblk):
            x = vss_block(x, sd, f"{pre}.layers.{i}.blocks.{j}")
        outs.append(ln(x, sd, f"{pre}.outnorm{i}").permute(0, 3, 1, 2).contiguous())
        if i < 3:
            x = patch_merging(x, sd, f"{pre}.layers.{i}.downsample")
    return outs


def cromb(x_rgb, x_e, sd, pre):
    """CrossMambaFusionBlock (vmamba.py:1857-1861) -> CrossMambaFusion_SS2D_SSM.forward (:1622-1640) ->
    Cross_Mamba_Attention_SSM.forward (:1508-1545).  pre = '...cross_mamba.i'."""
    op = pre + ".op"
    B, H, W, _ = x_rgb.shape
    D = sd[op + ".conv2d.weight"].shape[0]
    L = H * W

    def branch(x, wname):
        t = F.linear(x, sd[op + wname]).permute(0, 3, 1, 2).contiguous()
        t = F.silu(F.conv2d(t, sd[op + ".conv2d.weight"], sd[op + ".conv2d.bias"], padding=1, groups=D))  # one conv for both (:1629-1630)
        return t.flatten(2)  # (B, D, L)

    xr, xe = branch(x_rgb, ".in_proj.weight"), branch(x_e, ".in_proj_modalx.weight")
    cm = op + ".CMA_ssm"
    R = sd[cm + ".dt_proj_1.weight"].shape[1]
    N = sd[cm + ".A_log_1"].shape[1]

    def proj(x, i):
        dbl = F.linear(x.transpose(1, 2), sd[f"{cm}.x_proj_{i}.weight"])  # (B, L, R+2N)
        dt, Bm, Cm = torch.split(dbl, [R, N, N], dim=-1)
        dt = (dt @ sd[f"{cm}.dt_proj_{i}.weight"].t()).transpose(1, 2).contiguous()  # (B, D, L)
        return dt, Bm.transpose(1, 2).contiguous(), Cm.transpose(1, 2).contiguous()

    dt_r, B_r, C_r = proj(xr, 1)
    dt_e, B_e, C_e = proj(xe, 2)
    y_r = selective_scan(xr, dt_r, -torch.exp(sd[cm + ".A_log_1"].float()), B_r, C_e, sd[cm + ".D_1"].float(),
                         sd[cm + ".dt_proj_1.bias"].float(), True)  # C swapped (:1530)
    y_e = selective_scan(xe, dt_e, -torch.exp(sd[cm + ".A_log_2"].float()), B_e, C_r, sd[cm + ".D_2"].float(),
                         sd[cm + ".dt_proj_2.bias"].float(), True)  # (:1536)
    y_r = ln(y_r.transpose(1, 2), sd, cm + ".out_norm_1").view(B, H, W, D)
    y_e = ln(y_e.transpose(1, 2), sd, cm + ".out_norm_2").view(B, H, W, D)
    return (x_rgb + F.linear(y_r, sd[op + ".out_proj_rgb.weight"]),
            x_e + F.linear(y_e, sd[op + ".out_proj_e.weight"]))


def conmb(x_rgb, x_e, sd, pre):
    """ConcatMambaFusionBlock (vmamba.py:1915-1916) -> ConMB_SS2D.forward (:1265-1284) ->
    cross_selective_scan_multimodal_k2 (:369-430).  pre = '...channel_attn_mamba.i'."""
    op = pre + ".op"
    B, H, W, _ = x_rgb.shape
    D = sd[op + ".conv2d.weight"].shape[0]
    L = H * W
    tr = F.linear(x_rgb, sd[op + ".in_proj.weight"]).permute(0, 3, 1, 2).contiguous()
    te = F.linear(x_e, sd[op + ".in_proj_modalx.weight"]).permute(0, 3, 1, 2).contiguous()
    cr = F.silu(F.conv2d(tr, sd[op + ".conv2d.weight"], sd[op + ".conv2d.bias"], padding=1, groups=D))
    ce = F.silu(F.conv2d(te, sd[op + ".conv2d_modalx.weight"], sd[op + ".conv2d_modalx.bias"], padding=1, groups=D))
    seq = torch.cat([cr.flatten(2), ce.flatten(2)], dim=2)                     # (:130)
    xs = torch.stack([seq, seq.flip(-1)], dim=1)                               # (:131)
    ys = _core(xs, sd[op + ".x_proj_weight"], sd[op + ".dt_projs_weight"], sd[op + ".dt_projs_bias"],
               sd[op + ".A_logs"], sd[op + ".Ds"])
    y = ys[:, 0] + ys[:, 1].flip(-1)                                           # (:149)
    y_r = ln(y[:, :, :L].transpose(1, 2).contiguous().view(B, H, W, D), sd, op + ".out_norm1")
    y_e = ln(y[:, :, L:].transpose(1, 2).contiguous().view(B, H, W, D), sd, op + ".out_norm2")

    def se(t, fc):  # (:1276-1279): avg-pool of the PRE-conv projection -> Linear, SiLU, Linear, Sigmoid
        s = t.mean(dim=(2, 3))
        return torch.sigmoid(F.linear(F.silu(F.linear(s, sd[f"{op}.{fc}.0.weight"])), sd[f"{op}.{fc}.2.weight"]))

    g_r, g_e = se(tr, "fc1"), se(te, "fc2")
    y_r = y_r * g_e[:, None, None, :]                                          # cross-applied (:1280-1281)
    y_e = y_e * g_r[:, None, None, :]
    out = F.linear(torch.cat([y_r, y_e], dim=-1), sd[op + ".out_proj.weight"])
    return x_rgb + x_e + out


def rgbx_encoder(rgb, mx, sd, pre="backbone"):
    """RGBXTransformer.forward_features (dual_vmamba.py:78-107)."""
    o_r = backbone(rgb, sd, pre + ".vssm")
    o_x = backbone(mx, sd, pre + ".vssm")
    fused = []
    for i in range(4):
        cr, cx = cromb(o_r[i].permute(0, 2, 3, 1).contiguous(), o_x[i].permute(0, 2, 3, 1).contiguous(),
                       sd, f"{pre}.cross_mamba.{i}")
        fused.append(conmb(cr, cx, sd, f"{pre}.channel_attn_mamba.{i}").permute(0, 3, 1, 2).contiguous())
    return fused


def channel_attention_block(x, sd, pre):
    """ChannelAttentionBlock / ChannelAttention (vmamba.py:1725-1757); x is NCHW."""
    t = F.conv2d(x, sd[pre + ".cab.0.weight"], sd[pre + ".cab.0.bias"], padding=1)
    t = F.gelu(t)
    t = F.conv2d(t, sd[pre + ".cab.2.weight"], sd[pre + ".cab.2.bias"], padding=1)

    def fc(v):
        v = F.silu(F.conv2d(v, sd[pre + ".cab.3.fc.0.weight"]))
        return F.conv2d(v, sd[pre + ".cab.3.fc.2.weight"])

    attn = fc(F.adaptive_avg_pool2d(t, 1)) + fc(F.adaptive_max_pool2d(t, 1))
    return t * torch.sigmoid(attn)


def cvss_decoder_block(x, sd, pre):
    """CVSSDecoderBlock._forward (vmamba.py:1800-1805)."""
    x = x * sd[pre + ".scale1"] + ss2d(ln(x, sd, pre + ".norm1"), sd, pre + ".op")
    y = channel_attention_block(ln(x, sd, pre + ".norm2").permute(0, 3, 1, 2).contiguous(), sd, pre + ".conv_blk")
    y = y + (x * sd[pre + ".scale2"]).permute(0, 3, 1, 2)
    return y.permute(0, 2, 3, 1).contiguous()


def _bilinear(x_nhwc, size=None, scale=None):
    t = x_nhwc.permute(0, 3, 1, 2).contiguous()
    t = F.interpolate(t, size=size, scale_factor=scale, mode="bilinear", align_corners=False)
    return t.permute(0, 2, 3, 1).contiguous()


def mamba_decoder(feats, sd, pre="decode_head"):
    """MambaDecoder.forward (MambaDecoder.py:259-280, forward_up_features :222-239, up_x4 :272-280)."""
    x = feats[3].permute(0, 2, 3, 1).contiguous()
    # layers_up.0 = PatchExpand (MambaDecoder.py:12-30)
    x = F.linear(x, sd[pre + ".layers_up.0.expand.weight"])
    B, H, W, C = x.shape
    x = x.view(B, H, W, 2, 2, C // 4).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, C // 4)
    y = ln(x, sd, pre + ".layers_up.0.norm")
    for inx in (1, 2, 3):
        skip = feats[3 - inx]
        y = _bilinear(y, size=skip.shape[2:])                       # :231-232
        x = y + skip.permute(0, 2, 3, 1)
        lp = f"{pre}.layers_up.{inx}"
        nblk = _count(sd, re.escape(lp + ".blocks.") + r"(\d+)\.norm1\.weight")
        for j in range(nblk):
            x = cvss_decoder_block(x, sd, f"{lp}.blocks.{j}")
        if (lp + ".upsample.linear.weight") in sd:                  # UpsampleExpand (MambaDecoder.py:33-51)
            x = _bilinear(F.linear(x, sd[lp + ".upsample.linear.weight"]), scale=2)
            x = ln(x, sd, lp + ".upsample.norm")
        y = x
    x = ln(y, sd, pre + ".norm_up")
    # FinalUpsample_X4 (MambaDecoder.py:76-97)
    x = _bilinear(F.linear(x, sd[pre + ".up.linear1.weight"]), scale=2)
    x = _bilinear(F.linear(x, sd[pre + ".up.linear2.weight"]), scale=2)
    x = ln(x, sd, pre + ".up.norm")
    return F.conv2d(x.permute(0, 3, 1, 2).contiguous(), sd[pre + ".output.weight"])


def encoder_decoder(rgb, mx, sd):
    """EncoderDecoder.encode_decode (builder.py:128-139), deep_supervision=False, no aux head."""
    out = mamba_decoder(rgbx_encoder(rgb, mx, sd, "backbone"), sd, "decode_head")
    return F.interpolate(out, size=rgb.shape[2:], mode="bilinear", align_corners=False)


def hist_info(n_cl, pred, gt):
    """utils/metric.py:8-15 restated: confusion matrix (rows = ground truth, columns = prediction) over the pixels whose
    label is in [0, n_cl), the number of such pixels, and the number of correct ones."""
    pred, gt = np.asarray(pred), np.asarray(gt)
    k = (gt >= 0) & (gt < n_cl)
    hist = np.bincount(n_cl * gt[k].astype(np.int64) + pred[k].astype(np.int64), minlength=n_cl ** 2).reshape(n_cl, n_cl)
    return hist, int(k.sum()), int((pred[k] == gt[k]).sum())


def compute_score(hist, correct, labeled):
    """utils/metric.py:17-33 restated: per-class IoU, mIoU, frequency-weighted IoU, mean class accuracy, pixel accuracy."""
    hist = np.asarray(hist, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
        freq = hist.sum(1) / hist.sum()
        class_acc = np.diag(hist) / hist.sum(axis=1)
        return iou, float(np.nanmean(iou)), float((iou[freq > 0] * freq[freq > 0]).sum()), float(np.nanmean(class_acc)), correct / labeled


def mean_iou(pred, gt, n_cl):
    """utils/metric.py:8-29 (hist_info + compute_score), returns (iou per class, mIoU)."""
    k = (gt >= 0) & (gt < n_cl)
    hist = np.bincount(n_cl * gt[k].astype(int) + pred[k].astype(int), minlength=n_cl ** 2).reshape(n_cl, n_cl)
    iou = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
    return iou, float(np.nanmean(iou))


def selective_scan_torch(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False):
    """Differentiable pure-torch restatement of selective_scan_ref (selective_scan_interface.py:86-131), used to
    check the TRAINING path (autograd through sigma_b200.ops.SelectiveScan) on small shapes.  B, C: (b, G, N, L)."""
    b, d, L = u.shape
    G, N = B.shape[1], B.shape[2]
    delta = delta + (delta_bias[None, :, None] if delta_bias is not None else 0.0)
    if delta_softplus:
        delta = F.softplus(delta)
    Bx = B.repeat_interleave(d // G, dim=1)          # (b, d, N, L)
    Cx = C.repeat_interleave(d // G, dim=1)
    dA = torch.exp(delta[:, :, None, :] * A[None, :, :, None])
    dBu = (delta * u)[:, :, None, :] * Bx
    h = u.new_zeros((b, d, N))
    ys = []
    for l in range(L):
        h = dA[..., l] * h + dBu[..., l]
        ys.append((h * Cx[..., l]).sum(-1))
    y = torch.stack(ys, dim=-1)
    return y if D is None else y + u * D[None, :, None]
