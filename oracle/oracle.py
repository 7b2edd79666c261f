"""CPU oracle for the OnePose++ 2D-3D matcher hot path — TEST INFRASTRUCTURE ONLY.

A functional, fp32, plain-PyTorch-CPU restatement of ``OnePosePlus_model.forward``
(reference: src/models/OnePosePlus/OnePosePlusModel.py:96-201) working directly on a
``state_dict`` with the reference's 195 key names.  It exists to check the CUDA path; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``onepose_plus_plus_b200``) never does.

Pinning: the reference ships no golden vectors or tests for this path.  The
oracle is pinned against the reference's own code imported in the build container
(``oracle/ref_shims.py`` + ``tests/test_oracle_vs_reference.py``) and against fixtures generated
from that import (``oracle/make_golden.py`` -> ``tests/golden/``).  One assumption cannot be
verified offline: kornia 0.4.1's ``create_meshgrid`` / ``dsnt.spatial_expectation2d`` (not in
this image) are restated from their published definitions (x fastest, (x, y) order).

Every function cites the reference lines it follows (paths relative to the reference repo).
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CONFIG = {
    # configs/experiment/inference_onepose.yaml:26-109
    "loftr_backbone": {
        "type": "ResNetFPN",
        "resolution": [8, 2],
        "resnetfpn": {"block_type": "BasicBlock", "initial_dim": 128,
                      "block_dims": [128, 196, 256], "output_layers": [3, 1]},
        "pretrained": None,
        "pretrained_fix": False,
    },
    "interpol_type": "bilinear",
    "keypoints_encoding": {"enable": True, "type": "mlp_linear", "descriptor_dim": 256,
                           "keypoints_encoder": [32, 64, 128], "norm_method": "instancenorm"},
    "positional_encoding": {"enable": True, "pos_emb_shape": [256, 256]},
    "loftr_coarse": {"type": "LoFTR", "d_model": 256, "d_ffm": 128, "nhead": 8,
                     "layer_names": ["self", "cross"], "layer_iter_n": 3, "dropout": 0.0,
                     "attention": "linear", "norm_method": "layernorm", "kernel_fn": "elu + 1",
                     "d_kernel": 16, "redraw_interval": 2, "rezero": None, "final_proj": False},
    "coarse_matching": {"type": "dual-softmax", "thr": 0.1, "feat_norm_method": "sqrt_feat_dim",
                        "border_rm": 2, "dual_softmax": {"temperature": 0.08},
                        "train": {"train_padding": True, "train_coarse_percent": 0.3,
                                  "train_pad_num_gt_min": 200}},
    "loftr_fine": {"enable": True, "window_size": 5, "coarse_layer_norm": False, "type": "LoFTR",
                   "d_model": 128, "nhead": 8, "layer_names": ["self", "cross"],
                   "layer_iter_n": 1, "dropout": 0.0, "attention": "linear",
                   "norm_method": "layernorm", "kernel_fn": "elu + 1", "d_kernel": 16,
                   "redraw_interval": 2, "rezero": None, "final_proj": False},
    "fine_matching": {"enable": True, "type": "s2d", "s2d": {"type": "heatmap"}},
}


# ---------------------------------------------------------------------------------------------
# backbone  (backbone/resnet.py:20-45, 85-164)
# ---------------------------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, 1e-5)


def _basic_block(sd, p, x, stride):
    # resnet.py:36-45
    y = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)))
    y = _bn(sd, p + ".bn2", F.conv2d(y, sd[p + ".conv2.weight"], None, 1, 1))
    if stride != 1:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride))
    return F.relu(x + y)


def _fpn_head(sd, p, x):
    # resnet.py:112-124: conv3x3, BN, LeakyReLU(0.01), conv3x3
    x = F.conv2d(x, sd[p + ".0.weight"], None, 1, 1)
    x = F.leaky_relu(_bn(sd, p + ".1", x), 0.01)
    return F.conv2d(x, sd[p + ".3.weight"], None, 1, 1)


def backbone(sd, image, prefix="backbone."):
    """resnet.py:141-164 with output_layers [3, 1] -> (x3_out [B,256,H/8,W/8], x1_out [B,128,H/2,W/2])."""
    p = prefix
    x0 = F.relu(_bn(sd, p + "bn1", F.conv2d(image, sd[p + "conv1.weight"], None, 2, 3)))
    x1 = _basic_block(sd, p + "layer1.1", _basic_block(sd, p + "layer1.0", x0, 1), 1)
    x2 = _basic_block(sd, p + "layer2.1", _basic_block(sd, p + "layer2.0", x1, 2), 1)
    x3 = _basic_block(sd, p + "layer3.1", _basic_block(sd, p + "layer3.0", x2, 2), 1)
    x3_out = F.conv2d(x3, sd[p + "layer3_outconv.weight"])
    up = F.interpolate(x3_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x2_out = _fpn_head(sd, p + "layer2_outconv2", F.conv2d(x2, sd[p + "layer2_outconv.weight"]) + up)
    up = F.interpolate(x2_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x1_out = _fpn_head(sd, p + "layer1_outconv2", F.conv2d(x1, sd[p + "layer1_outconv.weight"]) + up)
    return x3_out, x1_out


# ---------------------------------------------------------------------------------------------
# encodings
# ---------------------------------------------------------------------------------------------
def position_encoding_sine(d_model, h, w):
    """position_encoding.py:13-35 including the floor-division quirk: the exponent scale is
    (-ln(1e4) / d_model // 2) == -1.0 for d_model 256, positions are 1-based."""
    y_pos = torch.arange(1, h + 1, dtype=torch.float32).view(1, h, 1).expand(1, h, w)
    x_pos = torch.arange(1, w + 1, dtype=torch.float32).view(1, 1, w).expand(1, h, w)
    div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))
    div = div[:, None, None]
    pe = torch.zeros(d_model, h, w)
    pe[0::4] = torch.sin(x_pos * div)
    pe[1::4] = torch.cos(x_pos * div)
    pe[2::4] = torch.sin(y_pos * div)
    pe[3::4] = torch.cos(y_pos * div)
    return pe


def normalize_3d_keypoints(kpts):
    """normalize.py:16-26: extents from batch element 0, per-batch mean centre, 0.6 * max extent."""
    ext = kpts[0].max(0).values - kpts[0].min(0).values
    center = kpts.mean(-2)
    return (kpts - center[:, None]) / (ext.max() * 0.6)


def keypoint_encoding(sd, kpts, descriptors, prefix="kpt_3d_pos_encoding.encoder."):
    """position_encoding.py:54-79, norm_method 'instancenorm': InstanceNorm1d applied to [B,N,C]
    normalises every point over its C features (biased variance, eps 1e-5, no affine)."""
    x = kpts
    for idx in (0, 3, 6, 9):
        x = F.linear(x, sd[f"{prefix}{idx}.weight"], sd[f"{prefix}{idx}.bias"])
        if idx != 9:
            m = x.mean(-1, keepdim=True)
            v = x.var(-1, unbiased=False, keepdim=True)
            x = F.relu((x - m) / torch.sqrt(v + 1e-5))
    return descriptors + x.transpose(2, 1)


# ---------------------------------------------------------------------------------------------
# transformer  (loftr_module/linear_attention.py:29-61, transformer.py:65-94, 133-171)
# ---------------------------------------------------------------------------------------------
def linear_attention(q, k, v, eps=1e-6, q_mask=None, kv_mask=None):
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    if q_mask is not None:      # linear_attention.py:49-53: padded positions are zeroed
        Q = Q * q_mask[:, :, None, None]
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None]
        v = v * kv_mask[:, :, None, None]
    v_len = v.size(1)
    v = v / v_len
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(1)) + eps)
    return torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * v_len


def full_attention(q, k, v):
    """FullAttention.forward (linear_attention.py:64-95), no masks, no dropout"""
    qk = torch.einsum("nlhd,nshd->nlsh", q, k)
    a = torch.softmax(qk / q.size(3) ** 0.5, dim=2)
    return torch.einsum("nlsh,nshd->nlhd", a, v).contiguous()


def encoder_layer(sd, p, x, source, nhead, x_mask=None, source_mask=None, attention="linear"):
    bs, d = x.size(0), x.size(2)
    dim = d // nhead
    q = F.linear(x, sd[p + "q_proj.weight"]).view(bs, -1, nhead, dim)
    k = F.linear(source, sd[p + "k_proj.weight"]).view(bs, -1, nhead, dim)
    v = F.linear(source, sd[p + "v_proj.weight"]).view(bs, -1, nhead, dim)
    if attention == "full":
        msg = full_attention(q, k, v).reshape(bs, -1, d)
    else:
        msg = linear_attention(q, k, v, q_mask=x_mask, kv_mask=source_mask).reshape(bs, -1, d)
    msg = F.layer_norm(F.linear(msg, sd[p + "merge.weight"]), (d,), sd[p + "norm1.weight"],
                       sd[p + "norm1.bias"], 1e-5)
    msg = F.linear(F.relu(F.linear(torch.cat([x, msg], 2), sd[p + "mlp.0.weight"])),
                   sd[p + "mlp.2.weight"])
    msg = F.layer_norm(msg, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    return x + msg


def local_feature_transformer(sd, prefix, cfg, desc3d, desc2d, collect=None, query_mask=None):
    """transformer.py:133-171.  desc3d [B,C,L] -> [B,L,C]; cross layers update both sequences from
    the PRE-update tensors (transformer.py:154-159); query_mask [B, S] masks the 2D side only
    (:150-159: x_mask and source_mask of the 2D self layer, x_mask of 2D<-3D, source_mask of 3D<-2D)."""
    names = list(cfg["layer_names"]) * cfg["layer_iter_n"]
    att = cfg.get("attention", "linear")
    d3 = desc3d.transpose(1, 2)
    d2 = desc2d
    for i, name in enumerate(names):
        p = f"{prefix}layers.{i}."
        if name == "self":
            d2, d3 = (encoder_layer(sd, p, d2, d2, cfg["nhead"], query_mask, query_mask, att),
                      encoder_layer(sd, p, d3, d3, cfg["nhead"], attention=att))
        elif name == "cross":
            d2, d3 = (encoder_layer(sd, p, d2, d3, cfg["nhead"], x_mask=query_mask, attention=att),
                      encoder_layer(sd, p, d3, d2, cfg["nhead"], source_mask=query_mask, attention=att))
        else:
            raise NotImplementedError(name)
        if collect is not None:
            collect.append((d3, d2))
    return d3, d2


# ---------------------------------------------------------------------------------------------
# coarse matching  (utils/coarse_matching.py:76-123, 125-242; inference branch)
# ---------------------------------------------------------------------------------------------
def coarse_matching(cfg, feat3d, feat2d, data, mask_query=None):
    c = feat3d.size(2)
    a = feat3d / c ** 0.5
    b = feat2d / c ** 0.5
    sim = torch.einsum("nlc,nsc->nls", a, b) / (cfg["dual_softmax"]["temperature"] + 1e-4)
    if mask_query is not None:   # coarse_matching.py:108-114: -1e9 added on the masked query cells
        neg = torch.zeros_like(sim)
        neg[~mask_query.bool()[:, None].expand_as(sim)] = -1e9
        sim = sim + neg
    conf = F.softmax(sim, 1) * F.softmax(sim, 2)
    data["conf_matrix"] = conf
    hc, wc = data["q_hw_c"]
    B, L, S = conf.shape
    mask = (conf > cfg["thr"]).view(B, L, hc, wc).clone()
    bd = cfg["border_rm"]
    # mask_border (coarse_matching.py:10-20): the `-b:0` slices are empty, only top/left cleared
    mask[:, :, :bd] = False
    mask[:, :, :, :bd] = False
    mask = mask.view(B, L, S)
    mask = mask * (conf == conf.max(2, keepdim=True)[0]) * (conf == conf.max(1, keepdim=True)[0])
    mask_v, all_j = mask.max(2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    scale = data["q_hw_i"][0] / hc
    if "query_image_scale" in data:
        scale_total = scale * data["query_image_scale"][b_ids][:, [1, 0]]
    else:
        scale_total = scale
    mkpts_query = torch.stack([j_ids % wc, j_ids // wc], 1) * scale_total
    keep = mconf != 0
    data.update({
        "b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids, "gt_mask": mconf == 0,
        "m_bids": b_ids[keep], "mkpts_3d_db": data["keypoints3d"][b_ids, i_ids][keep],
        "mkpts_query_c": mkpts_query[keep], "mconf": mconf[keep],
    })


# ---------------------------------------------------------------------------------------------
# fine level  (loftr_module/fine_preprocess.py:32-55, utils/fine_matching.py:28-110)
# ---------------------------------------------------------------------------------------------
def fine_preprocess(W, d_fine, data, desc3d_db, feat_f):
    data["W"] = W
    if data["b_ids"].shape[0] == 0:
        return torch.empty(0, d_fine, 1), torch.empty(0, W * W, d_fine)
    stride = data["q_hw_f"][0] // data["q_hw_c"][0]
    unf = F.unfold(feat_f, kernel_size=(W, W), stride=stride, padding=W // 2)
    n, cww, l = unf.shape
    unf = unf.view(n, cww // (W * W), W * W, l).permute(0, 3, 2, 1)  # 'n (c ww) l -> n l ww c'
    f3d = desc3d_db.permute(0, 2, 1)[data["b_ids"], data["i_ids"], :].unsqueeze(-1)
    return f3d, unf[data["b_ids"], data["j_ids"]]


def fine_matching(feat3d, feat2d_unfold, data):
    M, WW, C = feat2d_unfold.shape
    W = int(math.sqrt(WW))
    scale = data["q_hw_i"][0] / data["q_hw_f"][0]
    if M == 0:
        data.update({"expec_f": torch.empty(0, 3), "mkpts_query_f": data["mkpts_query_c"]})
        return
    f0 = feat3d[:, feat3d.shape[1] // 2, :]
    sim = torch.einsum("mc,mrc->mr", f0, feat2d_unfold)
    heat = torch.softmax(sim / C ** 0.5, 1)
    # kornia 0.4.1 create_meshgrid(W, W, normalized=True): linspace(-1, 1, W), (x, y) last, x fastest
    lin = torch.linspace(-1, 1, W)
    grid = torch.stack([lin.repeat(W), lin.repeat_interleave(W)], 1)  # [WW, 2]
    coords = heat @ grid  # dsnt.spatial_expectation2d
    var = heat @ grid ** 2 - coords ** 2
    std = torch.sqrt(torch.clamp(var, min=1e-10)).sum(-1)
    data["expec_f"] = torch.cat([coords, std[:, None]], -1)
    if "query_image_scale" in data:
        qscale = scale * data["query_image_scale"][data["b_ids"]][:, [1, 0]]
    else:
        qscale = scale
    data["mkpts_query_f"] = data["mkpts_query_c"] + (coords * (W // 2) * qscale)[: len(data["mkpts_query_c"])]


# ---------------------------------------------------------------------------------------------
# top level  (OnePosePlusModel.py:96-201)
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def forward(sd, data, cfg=DEFAULT_CONFIG, stages=None):
    """Mutates `data` like the reference.  `stages` (optional dict) receives intermediate tensors."""
    img = data["query_image"]
    data["bs"] = img.size(0)
    data["q_hw_i"] = img.shape[2:]
    feat_c, feat_f = backbone(sd, img)
    data["q_hw_c"] = feat_c.shape[2:]
    data["q_hw_f"] = feat_f.shape[2:]
    h, w = feat_c.shape[2:]
    pe = position_encoding_sine(cfg["loftr_coarse"]["d_model"], h, w)
    q_c = (feat_c + pe[None]).flatten(2).transpose(1, 2)  # 'n c h w -> n (h w) c'
    kp = normalize_3d_keypoints(data["keypoints3d"])
    dsel = data["descriptors3d_coarse_db"] if "descriptors3d_coarse_db" in data else data["descriptors3d_db"]
    d3 = keypoint_encoding(sd, kp, dsel)
    if stages is not None:
        stages.update(feat_c=feat_c, feat_f=feat_f, tok2d_in=q_c, tok3d_in=d3.transpose(1, 2))
        stages["layers"] = []
    query_mask = data["query_image_mask"].flatten(-2) if "query_image_mask" in data else None
    d3, q_c = local_feature_transformer(sd, "loftr_coarse.", cfg["loftr_coarse"], d3, q_c,
                                        None if stages is None else stages["layers"], query_mask)
    coarse_matching(cfg["coarse_matching"], d3, q_c, data, mask_query=query_mask)
    W = cfg["loftr_fine"]["window_size"]
    f3d, f2d = fine_preprocess(W, cfg["loftr_fine"]["d_model"], data, data["descriptors3d_db"], feat_f)
    if f2d.size(0) != 0 and cfg["loftr_fine"]["enable"]:
        f3d, f2d = local_feature_transformer(sd, "loftr_fine.", cfg["loftr_fine"], f3d, f2d)
    else:
        f3d = f3d.transpose(1, 2)
    if stages is not None:
        stages.update(fine3d=f3d, fine2d=f2d)
    fine_matching(f3d, f2d, data)
    return data
