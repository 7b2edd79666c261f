"""Drop-in for ``LoFTR_for_OnePose_Plus`` (src/KeypointFreeSfM/loftr_for_sfm/loftr.py:16-167) — the
2D-2D detector-free matcher OnePose++ uses for its keypoint-free SfM mapping and as the detector
of demo.py — on the same sm_100a kernels as the 2D-3D matcher.

Same constructor (``config, enable_fine_matching=True``; config = the lower-cased dict of
``loftr_for_onepose_plus_cfg.py``), same state-dict keys (``backbone.*``, ``loftr_coarse.layers.N.*``,
``loftr_fine.layers.N.*``; ``pos_encoding.pe`` non-persistent), same in-place ``forward(data)``
contract: reads ``image0, image1`` (+ ``scale0, scale1``), writes ``bs, hw*_i/c/f, conf_matrix,
b_ids, i_ids, j_ids, gt_mask, m_bids, mkpts0_c, mkpts1_c, mconf, W, expec_f, mkpts0_f, mkpts1_f``.
What differs from the 2D-3D path (submodules/LoFTR/src/loftr): both sequences are image tokens,
8 coarse layers with SEQUENTIAL cross updates (transformer.py:96-97), temperature 0.1 / threshold
0.2 / border removal on all sides of both grids (coarse_matching.py), W x W windows (W = 9) from
both fine maps and the centre token of image 0's window correlated with image 1's
(fine_matching.py).

Built: inference with predicted coarse matches, images of equal size per call, linear attention,
``fine_concat_coarse_feat`` False (the shipped configuration).  Not built (raise): padding masks
(``mask0/mask1``), the coarse-matches-given branch (``mkpts0_c`` in data), ``extract_*_feature``
sampling, training mode.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .model import (LocalFeatureTransformer, ResNetFPN_8_2, _Engine)

__all__ = ["LoFTR_for_OnePose_Plus"]


class _PositionEncodingSine(nn.Module):
    """submodules/LoFTR/src/loftr/utils/position_encoding.py:11-36 (both frequency tables)"""

    def __init__(self, d_model, max_shape=(256, 256), temp_bug_fix=True):
        super().__init__()
        pe = torch.zeros((d_model, *max_shape))
        y_position = torch.ones(max_shape).cumsum(0).float().unsqueeze(0)
        x_position = torch.ones(max_shape).cumsum(1).float().unsqueeze(0)
        k = torch.arange(0, d_model // 2, 2).float()
        if temp_bug_fix:
            div_term = torch.exp(k * (-math.log(10000.0) / (d_model // 2)))
        else:   # the historical table (backward compatibility of released checkpoints)
            div_term = torch.exp(k * (-math.log(10000.0) / d_model // 2))
        div_term = div_term[:, None, None]
        pe[0::4], pe[1::4] = torch.sin(x_position * div_term), torch.cos(x_position * div_term)
        pe[2::4], pe[3::4] = torch.sin(y_position * div_term), torch.cos(y_position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0), persistent=False)


def _transformer(cfg):
    # LoFTR's LocalFeatureTransformer takes (d_model, nhead, layer_names, attention) only
    return LocalFeatureTransformer({"d_model": cfg["d_model"], "nhead": cfg["nhead"], "type": "LoFTR",
                                    "layer_names": list(cfg["layer_names"]), "layer_iter_n": 1,
                                    "attention": cfg["attention"], "norm_method": "layernorm", "rezero": None,
                                    "redraw_interval": None, "final_proj": False})


class LoFTR_for_OnePose_Plus(_Engine):
    def __init__(self, config, enable_fine_matching=True, precision=None):
        super().__init__()
        self.config = config
        self.enable_fine_matching = enable_fine_matching
        if config["backbone_type"] != "ResNetFPN" or tuple(config["resolution"]) != (8, 2):
            raise ValueError(f"LOFTR.BACKBONE_TYPE {config['backbone_type']} / resolution not supported.")
        rf = config["resnetfpn"]
        if rf["initial_dim"] != 128 or list(rf["block_dims"]) != [128, 196, 256]:
            raise NotImplementedError("backbone kernels are built for dims 128/[128,196,256]")
        if config["coarse"]["d_model"] != 256 or config["coarse"]["nhead"] != 8 or \
                config["fine"]["d_model"] != 128 or config["fine"]["nhead"] != 8:
            raise NotImplementedError("kernels are built for d_model 256/128, 8 heads")
        if config["coarse"]["attention"] != "linear" or config["fine"]["attention"] != "linear":
            raise NotImplementedError("the 2D-2D matcher is built for attention='linear'")
        if config["match_coarse"]["match_type"] != "dual_softmax":
            raise NotImplementedError("match_type 'sinkhorn' is not built")
        if config["fine_concat_coarse_feat"]:
            raise NotImplementedError("fine_concat_coarse_feat=True is not built (False in loftr_for_onepose_plus_cfg.py)")
        self.W = config["fine_window_size"]
        if self.W % 2 != 1 or not 3 <= self.W <= 9:
            raise NotImplementedError("fine window sizes 3, 5, 7, 9 are built")
        self._init_engine(precision, "linear")
        self.backbone = ResNetFPN_8_2({"block_type": "BasicBlock", "initial_dim": rf["initial_dim"],
                                       "block_dims": list(rf["block_dims"]), "output_layers": [3, 1]})
        self.pos_encoding = _PositionEncodingSine(config["coarse"]["d_model"],
                                                  temp_bug_fix=config["coarse"]["temp_bug_fix"])
        self.loftr_coarse = _transformer(config["coarse"])
        self.loftr_fine = _transformer(config["fine"])

    def _pe_module(self):
        return self.pos_encoding

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_plan"], st["_plan_sig"], st["_ws"], st["_sig_tensors"], st["_graphs"] = None, None, {}, None, {}
        return st

    # ------------------------------------------------------------------ stages
    def _coarse(self, t0, t1, B, S0, S1):
        """LoFTR LocalFeatureTransformer.forward (loftr_module/transformer.py:81-101): self layers on
        both images, then cross layers one after the other — feat1 reads the UPDATED feat0."""
        dev = t0.device
        f16 = torch.float16
        pl = 2 if self.split else 1
        cur0, cur1 = t0, t1
        for i, name in enumerate(self.loftr_coarse.layer_names):
            L = self._plan["coarse"][i]
            n0 = self._buf(f"lt0_{i % 2}", (B, S0, pl * 256), f16, dev)
            n1 = self._buf(f"lt1_{i % 2}", (B, S1, pl * 256), f16, dev)
            if name == "self":
                self._encoder_layer(L, "c2_", cur0, cur0, B, S0, S0, n0)
                self._encoder_layer(L, "c3_", cur1, cur1, B, S1, S1, n1)
            else:
                self._encoder_layer(L, "c2_", cur0, cur1, B, S0, S1, n0)
                self._encoder_layer(L, "c3_", cur1, n0, B, S1, S0, n1)
            cur0, cur1 = n0, n1
        return cur0, cur1

    def _fine_layer(self, L, x, src, G, T, out16, out32=None):
        """LoFTREncoderLayer.forward (d_model 128) on G groups of T tokens: x, src [G*T, pl*128]."""
        dev = x.device
        f16 = torch.float16
        split = self.split
        pl = 2 if split else 1
        rows = G * T
        q = self._buf("lf_q", (rows, pl * 128), f16, dev)
        kv = self._buf("lf_kv", (rows, pl * 256), f16, dev)
        att = self._buf("lf_att", (rows, pl * 128), f16, dev)
        msg = self._buf("lf_msg", (rows, pl * 128), f16, dev)
        h = self._buf("lf_h", (rows, pl * 256), f16, dev)
        ops.linear_act(x, None, L["wq"], q, rows, 2, 128, split)           # Q = elu(q_proj x) + 1
        ops.linear_act(src, None, L["wkv"], kv, rows, 2, 128, split)       # K' = elu(k_proj s) + 1 | V
        ops.seq_attention(q, kv, att, G, T, T, split)
        ops.linear_ln(att, None, L["merge16"], False, *L["n1"], 1, rows, split, out16=msg)
        ops.linear_act(x, msg, L["mlp0"], h, rows, 1, 256, split)
        ops.linear_ln(h, None, L["mlp2"], False, *L["n2"], 1, rows, split, resid=x, out16=out16, out32=out32)

    # ------------------------------------------------------------------ forward
    def forward(self, data, **kwargs):
        if self.training:
            raise NotImplementedError("LoFTR_for_OnePose_Plus (B200) is the inference matcher: call .eval()")
        if "mask0" in data or "mask1" in data:
            raise NotImplementedError("padding masks (mask0 / mask1) are not built")
        if "mkpts0_c" in data:
            raise NotImplementedError("the fine-only branch with given coarse matches (loftr.py:81-121) is not built")
        if kwargs.get("extract_coarse_feature") or kwargs.get("extract_fine_feature"):
            raise NotImplementedError("feature extraction at the matches (loftr.py:136-165) is not built")
        im0, im1 = data["image0"], data["image1"]
        if not (torch.is_tensor(im0) and im0.is_cuda and im1.is_cuda):
            raise RuntimeError("LoFTR_for_OnePose_Plus (B200) has no CPU path: move the model and data to a CUDA device")
        if im0.dim() != 4 or im0.shape[1] != 1 or im1.shape != im0.shape:
            raise ValueError(f"image0 / image1 must both be [B, 1, H, W] of one size, got {tuple(im0.shape)}, "
                             f"{tuple(im1.shape)} (differently sized pairs: one call per size)")
        B, _, H, W = im0.shape
        if H % 8 or W % 8 or H < 48 or W < 48:
            raise ValueError("image height/width must be multiples of 8 (>= 48)")
        for k in ("scale0", "scale1"):
            if k in data and tuple(data[k].shape) != (B, 2):
                raise ValueError(f"{k} must be [B, 2]")
        if ("scale0" in data) != ("scale1" in data):
            raise ValueError("scale0 and scale1 come together")
        with torch.no_grad(), torch.cuda.device(im0.device):
            dev = im0.device
            self._ensure_plan(dev)
            split = self.split
            pl = 2 if split else 1
            f16, f32, i32 = torch.float16, torch.float32, torch.int32
            img = torch.cat([im0, im1], 0)
            if img.dtype not in (torch.uint8, torch.float32):
                img = img.float()
            tok, fine_map, (hc, wc) = self._backbone(img.contiguous())     # loftr.py:46-49 (one batched pass)
            S = hc * wc
            hf, wf = fine_map.shape[1:3]
            data.update({"bs": B, "hw0_i": im0.shape[2:], "hw1_i": im1.shape[2:],
                         "hw0_c": torch.Size((hc, wc)), "hw1_c": torch.Size((hc, wc)),
                         "hw0_f": torch.Size((hf, wf)), "hw1_f": torch.Size((hf, wf))})
            t0, t1 = self._coarse(tok[:B], tok[B:], B, S, S)
            # ---- coarse matching (LoFTR utils/coarse_matching.py:74-107, 133-259)
            mc = self.config["match_coarse"]
            scale = 1.0 / (256.0 * mc["dsmax_temperature"])
            ts = ops.sim_tiles(S)
            pm, ps = self._buf("pm_pt", (B * S, ts), f32, dev), self._buf("ps_pt", (B * S, ts), f32, dev)
            lse0, lse1 = self._buf("lse_pt", (B, S), f32, dev), self._buf("lse_px", (B, S), f32, dev)
            groups = (S + 31) // 32
            ops.sim_lse_cols(t0, t1, B, S, S, 256, scale, pm, ps, lse0, self._buf("lse_col_m", (B, groups, S), f32, dev),
                             self._buf("lse_col_s", (B, groups, S), f32, dev), lse1, split)
            conf = torch.empty((B, S, S), dtype=f32, device=dev)
            pt_val, pt_idx = self._buf("pt_val", (B, S), f32, dev), self._buf("pt_idx", (B, S), i32, dev)
            colmax = self._buf("colmax", (B, S), i32, dev)
            ops.sim_conf_colmax(t0, t1, lse0, lse1, conf, B, S, S, 256, scale, pm, self._buf("pi_pt", (B * S, ts), i32, dev),
                                pt_val, pt_idx, colmax, split)
            cap = B * S
            count = self._buf("match_count", (1,), i32, dev)
            b_ids, i_ids, j_ids = (torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(3))
            mconf = torch.empty(cap, dtype=f32, device=dev)
            mk0, mk1 = torch.empty((cap, 2), dtype=f32, device=dev), torch.empty((cap, 2), dtype=f32, device=dev)
            s0 = data["scale0"].to(device=dev, dtype=f32).contiguous() if "scale0" in data else None
            s1 = data["scale1"].to(device=dev, dtype=f32).contiguous() if "scale1" in data else None
            ops.match_select_2d(pt_val, pt_idx, colmax, s0, s1, B, hc, wc, hc, wc, mc["thr"], mc["border_rm"],
                                float(H / hc), self._buf("match_scratch", ((cap + 1023) // 1024 + 2,), i32, dev),
                                b_ids, i_ids, j_ids, mconf, mk0, mk1, count)
            M = int(count.item())   # the one host sync (the reference syncs in torch.where)
            b_ids, i_ids, j_ids = b_ids[:M], i_ids[:M], j_ids[:M]
            data.update({"conf_matrix": conf, "b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids,
                         "gt_mask": torch.zeros(M, dtype=torch.bool, device=dev), "m_bids": b_ids,
                         "mkpts0_c": mk0[:M], "mkpts1_c": mk1[:M], "mconf": mconf[:M]})
            if not self.enable_fine_matching:
                data.update({"mkpts0_f": data["mkpts0_c"], "mkpts1_f": data["mkpts1_c"]})
                return
            # ---- fine level (fine_preprocess.py:30-59, transformer.py:81-101, fine_matching.py:17-74)
            data["W"] = self.W
            if M == 0:
                data.update({"expec_f": torch.empty(0, 3, device=dev), "mkpts0_f": data["mkpts0_c"],
                             "mkpts1_f": data["mkpts1_c"]})
                return
            WW = self.W * self.W
            rows = 2 * M * WW
            xa = self._buf("lf_xa", (rows, pl * 128), f16, dev)
            xb = self._buf("lf_xb", (rows, pl * 128), f16, dev)
            x32 = self._buf("lf_x32", (rows, 128), f32, dev)
            ops.fine_gather_2d(fine_map[:B], fine_map[B:], b_ids, i_ids, j_ids, xa, M, hf, wf, wc, hf, wf, wc,
                               hf // hc, self.W, split)
            half = M * WW
            cur, nxt = xa, xb
            names = self.loftr_fine.layer_names
            for li, name in enumerate(names):
                L = self._plan["fine"][li]
                last = li == len(names) - 1
                if name == "self":      # both windows at once: 2M groups attending to themselves
                    self._fine_layer(L, cur, cur, 2 * M, WW, nxt, x32 if last else None)
                else:                   # sequential: window 0 from window 1, then window 1 from the NEW window 0
                    self._fine_layer(L, cur[:half], cur[half:], M, WW, nxt[:half], x32[:half] if last else None)
                    self._fine_layer(L, cur[half:], nxt[:half], M, WW, nxt[half:], x32[half:] if last else None)
                cur, nxt = nxt, cur
            if not names:
                raise NotImplementedError("a fine transformer without layers is not built")
            expec_f = torch.empty((M, 3), dtype=f32, device=dev)
            mk1f = torch.empty((M, 2), dtype=f32, device=dev)
            ops.fine_match_2d(x32, data["mkpts1_c"], b_ids, s1, expec_f, mk1f, M, self.W, float(H / hf))
            data.update({"expec_f": expec_f, "mkpts0_f": data["mkpts0_c"], "mkpts1_f": mk1f})
