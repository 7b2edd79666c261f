"""Synthetic checkpoint + planted 2D-3D workload — TEST INFRASTRUCTURE ONLY.

The reference's pretrained weights and datasets are Google-Drive downloads (README.md:36,50) and
are not available offline, and with random weights + random descriptors the matcher returns zero
matches.  This module builds (i) a seeded synthetic ``state_dict`` with the
reference's 195 keys and (ii) a workload whose 3D descriptors are *planted* from the image's own
features so that the coarse stage finds real mutual matches and the fine stage has non-trivial
sub-pixel offsets (recipe: SURVEY.md App. B).
"""
import math

import torch

from . import oracle


def _kaiming_fan_out(g, shape):
    fan_out = shape[0] * shape[2] * shape[3]
    return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)


def _xavier(g, shape):
    bound = math.sqrt(6.0 / (shape[0] + shape[1]))
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def synthetic_loftr_state_dict(seed=0, gain=1.0):
    """Seeded weights with the layout of LoFTR_for_OnePose_Plus (loftr.py:24-33): the backbone keys
    of the 2D-3D matcher, 8 coarse layers, 2 fine layers.  `gain` scales the last coarse layer's
    norm2 affine: with default-init weights the coarse tokens of different cells are almost
    parallel and no confidence passes the 0.2 threshold; a synthetic checkpoint may as well be a
    contrasty one."""
    sd = {k: v for k, v in synthetic_state_dict(seed).items()
          if k.startswith("backbone.") or k.startswith("loftr_fine.")}
    g = torch.Generator().manual_seed(seed + 4242)
    for i in range(8):
        q = f"loftr_coarse.layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj", "merge"):
            sd[q + proj + ".weight"] = _xavier(g, (256, 256))
        sd[q + "mlp.0.weight"] = _xavier(g, (512, 512))
        sd[q + "mlp.2.weight"] = _xavier(g, (256, 512))
        for nrm in ("norm1", "norm2"):
            sd[q + nrm + ".weight"] = 1.0 + 0.05 * torch.randn(256, generator=g)
            sd[q + nrm + ".bias"] = 0.05 * torch.randn(256, generator=g)
    sd["loftr_coarse.layers.7.norm2.weight"] *= gain
    sd["loftr_coarse.layers.7.norm2.bias"] *= gain
    return sd


def loftr_pair(h=256, w=320, batch=1, shift=(16, 24), seed=1, noise=0.01, with_scale=False):
    """Image pair for the 2D-2D matcher: image1 is image0 translated by `shift` pixels (multiples
    of 8 keep the coarse cells aligned, the remainder gives sub-cell fine offsets) plus noise."""
    g = torch.Generator().manual_seed(seed)
    dy, dx = shift
    big = torch.rand(batch, 1, h + abs(dy), w + abs(dx), generator=g)
    im0 = big[:, :, :h, :w].contiguous()
    im1 = (big[:, :, dy:dy + h, dx:dx + w] + noise * torch.randn(batch, 1, h, w, generator=g)).clamp(0, 1).contiguous()
    data = {"image0": im0, "image1": im1}
    if with_scale:
        data["scale0"] = torch.tensor([[1.25, 0.8]]).expand(batch, -1).contiguous()
        data["scale1"] = torch.tensor([[0.9, 1.1]]).expand(batch, -1).contiguous()
    return data


@torch.no_grad()
def planted_loftr(h=256, w=320, batch=1, shift=(16, 24), seed=0, backbone_gain=4.0, layer_gain=0.3,
                  with_scale=False, noise=0.01):
    """Checkpoint + image pair on which the 2D-2D matcher finds hundreds of geometrically consistent
    matches (i - j = the planted shift in coarse cells).  A default-init LoFTR collapses all coarse
    tokens onto one direction, so the synthetic checkpoint is made contrasty: the coarse layers'
    second LayerNorm is damped (`layer_gain`), the coarse backbone head amplified (`backbone_gain`),
    and minus the mean final token is folded into the last layer's norm2.bias (the dual softmax is
    only discriminative on centred tokens).  Returns (state_dict, data)."""
    from . import loftr_oracle
    data = loftr_pair(h, w, batch=batch, shift=shift, seed=seed + 1, noise=noise, with_scale=with_scale)
    sd = synthetic_loftr_state_dict(seed)
    for i in range(8):
        sd[f"loftr_coarse.layers.{i}.norm2.weight"] *= layer_gain
        sd[f"loftr_coarse.layers.{i}.norm2.bias"] *= layer_gain
    sd["backbone.layer3_outconv.weight"] = sd["backbone.layer3_outconv.weight"] * backbone_gain
    cfg = loftr_oracle.DEFAULT_CONFIG
    fc, _ = oracle.backbone(sd, torch.cat([data["image0"], data["image1"]], 0))
    pe = loftr_oracle.position_encoding(256, *fc.shape[2:], cfg["coarse"]["temp_bug_fix"])
    t = (fc + pe[None]).flatten(2).transpose(1, 2)
    t0, t1 = loftr_oracle.transformer(sd, "loftr_coarse.", cfg["coarse"], t[:batch], t[batch:])
    sd["loftr_coarse.layers.7.norm2.bias"] = sd["loftr_coarse.layers.7.norm2.bias"] - torch.cat([t0, t1], 1).mean((0, 1))
    # the amplified coarse head also flows down the FPN: bring the fine map back to unit scale, else
    # the fine correlation logits reach ~1e4 and amplify fp32 rounding itself to 1e-2
    _, ff = oracle.backbone(sd, data["image0"][:1])
    sd["backbone.layer1_outconv2.3.weight"] = sd["backbone.layer1_outconv2.3.weight"] * (3.0 / ff.std())
    return sd, data


def synthetic_state_dict(seed=0, bn_perturb=True):
    """Seeded weights with the reference layout.  Initialisers follow the
    reference (kaiming fan_out convs resnet.py:126-131, xavier transformer transformer.py:128-131);
    BatchNorm affine/running statistics are perturbed so that BN folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k):
        sd[name + ".weight"] = _kaiming_fan_out(g, (co, ci, k, k))

    def bn(name, c):
        if bn_perturb:
            sd[name + ".weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
            sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_var"] = 0.6 + 0.8 * torch.rand(c, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(c)
            sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c)
            sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def block(name, ci, co, stride):
        conv(name + ".conv1", co, ci, 3)
        conv(name + ".conv2", co, co, 3)
        bn(name + ".bn1", co)
        bn(name + ".bn2", co)
        if stride != 1:
            conv(name + ".downsample.0", co, ci, 1)
            bn(name + ".downsample.1", co)

    p = "backbone."
    conv(p + "conv1", 128, 1, 7)
    bn(p + "bn1", 128)
    dims = [128, 196, 256]
    cin = 128
    for li, (d, s) in enumerate(zip(dims, (1, 2, 2)), start=1):
        block(f"{p}layer{li}.0", cin, d, s)
        block(f"{p}layer{li}.1", d, d, 1)
        cin = d
    conv(p + "layer3_outconv", 256, 256, 1)
    conv(p + "layer2_outconv", 256, 196, 1)
    conv(p + "layer2_outconv2.0", 256, 256, 3)
    bn(p + "layer2_outconv2.1", 256)
    conv(p + "layer2_outconv2.3", 196, 256, 3)
    conv(p + "layer1_outconv", 196, 128, 1)
    conv(p + "layer1_outconv2.0", 196, 196, 3)
    bn(p + "layer1_outconv2.1", 196)
    conv(p + "layer1_outconv2.3", 128, 196, 3)

    chans = [3, 32, 64, 128, 256]
    for i, idx in enumerate((0, 3, 6, 9)):
        bound = 1.0 / math.sqrt(chans[i])
        sd[f"kpt_3d_pos_encoding.encoder.{idx}.weight"] = \
            (torch.rand(chans[i + 1], chans[i], generator=g) * 2 - 1) * bound
        sd[f"kpt_3d_pos_encoding.encoder.{idx}.bias"] = \
            torch.zeros(chans[i + 1]) if idx == 9 else (torch.rand(chans[i + 1], generator=g) * 2 - 1) * bound

    for name, d, n_layers in (("loftr_coarse", 256, 6), ("loftr_fine", 128, 2)):
        for i in range(n_layers):
            q = f"{name}.layers.{i}."
            for proj in ("q_proj", "k_proj", "v_proj", "merge"):
                sd[q + proj + ".weight"] = _xavier(g, (d, d))
            sd[q + "mlp.0.weight"] = _xavier(g, (2 * d, 2 * d))
            sd[q + "mlp.2.weight"] = _xavier(g, (d, 2 * d))
            for nrm in ("norm1", "norm2"):
                sd[q + nrm + ".weight"] = 1.0 + 0.05 * torch.randn(d, generator=g)
                sd[q + nrm + ".bias"] = 0.05 * torch.randn(d, generator=g)
    return sd


@torch.no_grad()
def planted_workload(sd, h=512, w=512, n_points=5000, n_planted=3000, batch=1, alpha=20.0,
                     beta=8.0, seed=1, image_noise=0.02, with_scale=True):
    """Inputs for OnePosePlus_model.forward whose bank is planted from image 0's own features.

    Images 1..batch-1 are image 0 plus a little noise (distinct inputs, matches still found);
    the bank is the same for every batch element (BASELINE.json configs 3/4: shared cloud).
    """
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 1, h, w, generator=g)
    imgs = [base]
    for _ in range(batch - 1):
        imgs.append((base + image_noise * torch.randn(1, 1, h, w, generator=g)).clamp(0, 1))
    image = torch.cat(imgs, 0)
    kpts = torch.rand(1, n_points, 3, generator=g) - 0.5

    fc, ff = oracle.backbone(sd, base)
    hc, wc = fc.shape[2:]
    pe = oracle.position_encoding_sine(256, hc, wc)
    qc = (fc + pe[None]).flatten(2).transpose(1, 2)[0]  # [S, 256]
    mu = qc.mean(0)
    r = qc - mu
    mu_dir = mu / mu.norm()
    r = r - (r @ mu_dir)[:, None] * mu_dir[None]
    ys, xs = torch.meshgrid(torch.arange(2, hc), torch.arange(2, wc), indexing="ij")
    interior = (ys * wc + xs).flatten()
    n_planted = min(n_planted, interior.numel(), n_points)
    cells = interior[torch.randperm(interior.numel(), generator=g)[:n_planted]]

    kenc = oracle.keypoint_encoding(sd, oracle.normalize_3d_keypoints(kpts),
                                    torch.zeros(1, 256, n_points))  # [1, 256, N]
    dc = torch.randn(1, 256, n_points, generator=g) * r.std() * alpha / 4
    dc[0, :, :n_planted] = alpha * r[cells].t() - kenc[0, :, :n_planted]

    hf, wf = ff.shape[2:]
    stride = hf // hc
    ffc = ff[0] - ff[0].mean((1, 2), keepdim=True)
    off = torch.randint(-2, 3, (n_planted, 2), generator=g)
    fy = ((cells // wc) * stride + off[:, 0]).clamp(0, hf - 1)
    fx = ((cells % wc) * stride + off[:, 1]).clamp(0, wf - 1)
    df = torch.randn(1, 128, n_points, generator=g) * ff.std()
    df[0, :, :n_planted] = beta * ffc[:, fy, fx]

    data = {
        "query_image": image,
        "keypoints3d": kpts.expand(batch, -1, -1).contiguous(),
        "descriptors3d_db": df.expand(batch, -1, -1).contiguous(),
        "descriptors3d_coarse_db": dc.expand(batch, -1, -1).contiguous(),
    }
    if with_scale:
        data["query_image_scale"] = torch.tensor([[1.25, 0.8]]).expand(batch, -1).contiguous()
    meta = {"cells": cells, "offsets": off, "n_planted": n_planted}
    return data, meta


@torch.no_grad()
def hetero_workload(sd, h=256, w=320, n_points=1500, n_planted=700, batch=3, seed=5):
    """Batch whose elements are DIFFERENT objects: every batch element has its own image, its own
    keypoint cloud (different extents and offsets), its own planted descriptor banks and its own
    query_image_scale — exercises the per-batch indexing (desc[b], kpts[b], img_scale[b]) and the
    reference quirk of per-batch mean with batch-0 extents (normalize.py:16-26), which a shared bank
    cannot see.  Each element is planted on its own image with the keypoint encoding it receives
    INSIDE the batch."""
    parts = [planted_workload(sd, h, w, n_points, n_planted, batch=1, seed=seed + 10 * b)[0] for b in range(batch)]
    g = torch.Generator().manual_seed(seed + 999)
    kpts = torch.cat([p["keypoints3d"] * (1.0 + 0.35 * b) + 0.2 * b for b, p in enumerate(parts)], 0)
    # re-plant the coarse descriptors with the encoding each element gets inside the batch
    kenc_batch = oracle.keypoint_encoding(sd, oracle.normalize_3d_keypoints(kpts), torch.zeros(batch, 256, n_points))
    dcs = []
    for b, p in enumerate(parts):
        kenc_alone = oracle.keypoint_encoding(sd, oracle.normalize_3d_keypoints(p["keypoints3d"]),
                                              torch.zeros(1, 256, n_points))
        dcs.append(p["descriptors3d_coarse_db"] + kenc_alone - kenc_batch[b:b + 1])
    scales = 0.7 + 0.8 * torch.rand(batch, 2, generator=g)
    return {
        "query_image": torch.cat([p["query_image"] for p in parts], 0),
        "keypoints3d": kpts.contiguous(),
        "descriptors3d_db": torch.cat([p["descriptors3d_db"] for p in parts], 0),
        "descriptors3d_coarse_db": torch.cat(dcs, 0),
        "query_image_scale": scales,
    }


def pad_mask(batch, hc, wc, seed=3):
    """query_image_mask as the img_pad data flow produces it (bool [B, hc, wc], False = padding):
    every image keeps a different top-left rectangle of valid coarse cells."""
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(batch, hc, wc, dtype=torch.bool)
    for b in range(batch):
        vh = hc - int(torch.randint(0, max(hc // 3, 1) + 1, (1,), generator=g))
        vw = wc - int(torch.randint(0, max(wc // 3, 1) + 1, (1,), generator=g))
        m[b, :vh, :vw] = True
    return m


def random_workload(h=512, w=512, n_points=5000, batch=1, seed=1):
    """BASELINE.json config 1 taken literally: random image, random descriptors (yields M = 0)."""
    g = torch.Generator().manual_seed(seed)
    return {
        "query_image": torch.rand(batch, 1, h, w, generator=g),
        "keypoints3d": (torch.rand(1, n_points, 3, generator=g) - 0.5).expand(batch, -1, -1).contiguous(),
        "descriptors3d_db": torch.randn(1, 128, n_points, generator=g).expand(batch, -1, -1).contiguous(),
        "descriptors3d_coarse_db": torch.randn(1, 256, n_points, generator=g).expand(batch, -1, -1).contiguous(),
    }
