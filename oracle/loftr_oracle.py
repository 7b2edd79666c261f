"""CPU oracle for the LoFTR 2D-2D matcher used by OnePose++'s mapping / detection stages
(`LoFTR_for_OnePose_Plus.forward`, src/KeypointFreeSfM/loftr_for_sfm/loftr.py:16-167, built from
submodules/LoFTR/src/loftr) — TEST INFRASTRUCTURE ONLY (see oracle/oracle.py's header).

Restated on a plain state_dict with the reference's key names (backbone.*, loftr_coarse.layers.N.*,
loftr_fine.layers.N.*).  Differences from the 2D-3D matcher that matter for parity:
  * both sequences are image tokens; cross layers are SEQUENTIAL: feat1 attends to the already
    updated feat0 (LoFTR transformer.py:96-97)
  * coarse matching: temperature 0.1 without the +1e-4, threshold 0.2, border of 2 cells removed
    on ALL four sides of BOTH grids (LoFTR coarse_matching.py:9-28,100-107,197-209)
  * fine level: W x W windows (W = 9 in loftr_for_onepose_plus_cfg.py:13) from both fine maps, the
    centre token of image 0's window against image 1's window (LoFTR fine_matching.py:46-70)
"""
import math

import torch
import torch.nn.functional as F

from . import oracle

DEFAULT_CONFIG = {
    # src/KeypointFreeSfM/loftr_for_sfm/utils/loftr_for_onepose_plus_cfg.py:10-46 (lower-cased)
    "backbone_type": "ResNetFPN", "resolution": (8, 2), "fine_window_size": 9, "fine_concat_coarse_feat": False,
    "resnetfpn": {"initial_dim": 128, "block_dims": [128, 196, 256]},
    "coarse": {"d_model": 256, "d_ffn": 256, "nhead": 8, "layer_names": ["self", "cross"] * 4,
               "attention": "linear", "temp_bug_fix": False},
    "match_coarse": {"thr": 0.2, "border_rm": 2, "match_type": "dual_softmax", "dsmax_temperature": 0.1,
                     "skh_iters": 3, "skh_init_bin_score": 1.0, "skh_prefilter": True,
                     "train_coarse_percent": 0.4, "train_pad_num_gt_min": 200},
    "fine": {"d_model": 128, "d_ffn": 128, "nhead": 8, "layer_names": ["self", "cross"], "attention": "linear"},
}


def position_encoding(d_model, h, w, temp_bug_fix):
    """LoFTR utils/position_encoding.py:13-36: the fixed and the historical (buggy) frequency tables"""
    if not temp_bug_fix:
        return oracle.position_encoding_sine(d_model, h, w)
    y_pos = torch.arange(1, h + 1, dtype=torch.float32).view(1, h, 1).expand(1, h, w)
    x_pos = torch.arange(1, w + 1, dtype=torch.float32).view(1, 1, w).expand(1, h, w)
    div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / (d_model // 2)))[:, None, None]
    pe = torch.zeros(d_model, h, w)
    pe[0::4], pe[1::4] = torch.sin(x_pos * div), torch.cos(x_pos * div)
    pe[2::4], pe[3::4] = torch.sin(y_pos * div), torch.cos(y_pos * div)
    return pe


def transformer(sd, prefix, cfg, f0, f1):
    """LoFTR loftr_module/transformer.py:81-101 (no masks)"""
    for i, name in enumerate(cfg["layer_names"]):
        p = f"{prefix}layers.{i}."
        if name == "self":
            f0 = oracle.encoder_layer(sd, p, f0, f0, cfg["nhead"])
            f1 = oracle.encoder_layer(sd, p, f1, f1, cfg["nhead"])
        else:
            f0 = oracle.encoder_layer(sd, p, f0, f1, cfg["nhead"])
            f1 = oracle.encoder_layer(sd, p, f1, f0, cfg["nhead"])   # the UPDATED feat0
    return f0, f1


def coarse_matching(cfg, fc0, fc1, data):
    """LoFTR utils/coarse_matching.py:74-107,133-259 (dual softmax, inference branch)"""
    c = fc0.shape[-1]
    sim = torch.einsum("nlc,nsc->nls", fc0 / c ** 0.5, fc1 / c ** 0.5) / cfg["dsmax_temperature"]
    conf = F.softmax(sim, 1) * F.softmax(sim, 2)
    data["conf_matrix"] = conf
    (h0, w0), (h1, w1) = data["hw0_c"], data["hw1_c"]
    B = conf.shape[0]
    mask = (conf > cfg["thr"]).view(B, h0, w0, h1, w1).clone()
    b = cfg["border_rm"]
    if b > 0:
        mask[:, :b] = False
        mask[:, :, :b] = False
        mask[:, :, :, :b] = False
        mask[:, :, :, :, :b] = False
        mask[:, -b:] = False
        mask[:, :, -b:] = False
        mask[:, :, :, -b:] = False
        mask[:, :, :, :, -b:] = False
    mask = mask.view(B, h0 * w0, h1 * w1)
    mask = mask * (conf == conf.max(2, keepdim=True)[0]) * (conf == conf.max(1, keepdim=True)[0])
    mask_v, all_j = mask.max(2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    scale = data["hw0_i"][0] / h0
    s0 = scale * data["scale0"][b_ids] if "scale0" in data else scale
    s1 = scale * data["scale1"][b_ids] if "scale1" in data else scale
    data.update({"b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids, "gt_mask": mconf == 0, "m_bids": b_ids,
                 "mkpts0_c": torch.stack([i_ids % w0, i_ids // w0], 1) * s0,
                 "mkpts1_c": torch.stack([j_ids % w1, j_ids // w1], 1) * s1, "mconf": mconf})


def fine_preprocess(W, d_f, data, ff0, ff1):
    """LoFTR loftr_module/fine_preprocess.py:30-59 with fine_concat_coarse_feat False"""
    data["W"] = W
    if data["b_ids"].shape[0] == 0:
        return torch.empty(0, W * W, d_f), torch.empty(0, W * W, d_f)
    stride = data["hw0_f"][0] // data["hw0_c"][0]

    def unfold(f):
        u = F.unfold(f, kernel_size=(W, W), stride=stride, padding=W // 2)
        n, cww, l = u.shape
        return u.view(n, cww // (W * W), W * W, l).permute(0, 3, 2, 1)

    return unfold(ff0)[data["b_ids"], data["i_ids"]], unfold(ff1)[data["b_ids"], data["j_ids"]]


def fine_matching(f0, f1, data):
    """LoFTR utils/fine_matching.py:17-74"""
    M, WW, C = f0.shape
    W = int(math.sqrt(WW))
    scale = data["hw0_i"][0] / data["hw0_f"][0]
    if M == 0:
        data.update({"expec_f": torch.empty(0, 3), "mkpts0_f": data["mkpts0_c"], "mkpts1_f": data["mkpts1_c"]})
        return
    heat = torch.softmax(torch.einsum("mc,mrc->mr", f0[:, WW // 2, :], f1) / C ** 0.5, 1)
    lin = torch.linspace(-1, 1, W)
    grid = torch.stack([lin.repeat(W), lin.repeat_interleave(W)], 1)
    coords = heat @ grid
    var = heat @ grid ** 2 - coords ** 2
    std = torch.sqrt(torch.clamp(var, min=1e-10)).sum(-1)
    data["expec_f"] = torch.cat([coords, std[:, None]], -1)
    s1 = scale * data["scale1"][data["b_ids"]] if "scale0" in data else scale
    data["mkpts0_f"] = data["mkpts0_c"]
    data["mkpts1_f"] = data["mkpts1_c"] + (coords * (W // 2) * s1)[: len(data["mconf"])]


@torch.no_grad()
def forward(sd, data, cfg=DEFAULT_CONFIG, enable_fine_matching=True):
    """LoFTR_for_OnePose_Plus.forward (loftr.py:35-127), coarse matches predicted (no 'mkpts0_c' input)"""
    im0, im1 = data["image0"], data["image1"]
    data.update({"bs": im0.size(0), "hw0_i": im0.shape[2:], "hw1_i": im1.shape[2:]})
    if data["hw0_i"] == data["hw1_i"]:
        fc, ff = oracle.backbone(sd, torch.cat([im0, im1], 0))
        (fc0, fc1), (ff0, ff1) = fc.split(data["bs"]), ff.split(data["bs"])
    else:
        (fc0, ff0), (fc1, ff1) = oracle.backbone(sd, im0), oracle.backbone(sd, im1)
    data.update({"hw0_c": fc0.shape[2:], "hw1_c": fc1.shape[2:], "hw0_f": ff0.shape[2:], "hw1_f": ff1.shape[2:]})
    d = cfg["coarse"]["d_model"]
    t0 = (fc0 + position_encoding(d, *fc0.shape[2:], cfg["coarse"]["temp_bug_fix"])[None]).flatten(2).transpose(1, 2)
    t1 = (fc1 + position_encoding(d, *fc1.shape[2:], cfg["coarse"]["temp_bug_fix"])[None]).flatten(2).transpose(1, 2)
    t0, t1 = transformer(sd, "loftr_coarse.", cfg["coarse"], t0, t1)
    coarse_matching(cfg["match_coarse"], t0, t1, data)
    if not enable_fine_matching:
        data.update({"mkpts0_f": data["mkpts0_c"], "mkpts1_f": data["mkpts1_c"]})
        return data
    u0, u1 = fine_preprocess(cfg["fine_window_size"], cfg["fine"]["d_model"], data, ff0, ff1)
    if u0.size(0) != 0:
        u0, u1 = transformer(sd, "loftr_fine.", cfg["fine"], u0, u1)
    fine_matching(u0, u1, data)
    return data
