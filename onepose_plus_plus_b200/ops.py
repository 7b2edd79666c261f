"""Thin tensor-level wrappers over the C ABI (include/opp_b200.h).  Each wrapper checks dtype /
contiguity, passes raw device pointers and the current CUDA stream, and raises on error.  There
is deliberately no alternative code path: without libopp_b200.so these functions raise."""
import torch

from . import _lib

ptr, call, stream = _lib.ptr, _lib.call, _lib.stream


def to_planes(x, split):
    """fp32 tensor [..., C] -> fp16 storage [..., planes*C]: (hi | lo) planes when split."""
    hi = x.half()
    if not split:
        return hi.contiguous()
    lo = (x - hi.float()).half()
    return torch.cat([hi, lo], -1).contiguous()


def from_planes(t, split):
    """Inverse of to_planes (returns fp32)."""
    if not split:
        return t.float()
    c = t.shape[-1] // 2
    return t[..., :c].float() + t[..., c:].float()


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def conv1_gemm(image, w64, a_buf, out, split):
    """conv1 7x7/2 + folded BN + ReLU as im2col + ONE 64-wide tcgen05 K chunk (bias rides in K column
    49).  image fp32 or uint8 [B,1,H,W]; w64 fp16 [C, planes*64]; a_buf fp16 [B*H/2*W/2, planes*64]."""
    B, _, H, W = image.shape
    if image.dtype not in (torch.float32, torch.uint8):
        raise TypeError(f"image: expected float32 or uint8, got {image.dtype}")
    call("opp_conv1_im2col", ptr(image), int(image.dtype == torch.uint8), ptr(a_buf), B, H, W, int(split),
         stream())
    rows = B * (H // 2) * (W // 2)
    call("opp_linear_act_f16", ptr(a_buf), 64, None, 0, ptr(w64), ptr(out), rows, w64.shape[0], 1,
         (w64.shape[0] + 31) // 32 * 32, int(split), stream())
    return out


def conv2d_nhwc(x, w, bias, out, ksize, stride, split, act=0, resid=None, slope=0.01, tok=None,
                pe=None, up=None):
    """x NHWC fp16 [B,H,W,planes*Cin_pad]; w fp16 [Cout_pad, planes*k*k*Cin_pad];
    act 0 none / 1 relu / 2 leaky; up: coarser NHWC map added as bilinear x2 (align_corners)."""
    _chk(x, torch.float16, "x")
    _chk(w, torch.float16, "w")
    _chk(resid, torch.float16, "resid")
    B, H, W, C = x.shape
    planes = 2 if split else 1
    call("opp_conv2d_nhwc", ptr(x), ptr(w), ptr(bias), ptr(resid), ptr(out), B, H, W, C // planes,
         w.shape[0], ksize, stride, act, float(slope), ptr(tok), ptr(pe), ptr(up), int(split), stream())
    return out


def kpt_encode(kpts, desc, mlp, stats, tok, split):
    B, N, _ = kpts.shape
    _chk(kpts, torch.float32, "keypoints3d")
    _chk(desc, torch.float32, "descriptors3d")
    call("opp_kpt_stats", ptr(kpts), ptr(stats), B, N, stream())
    (w1, b1), (w2, b2), (w3, b3), (w4, b4) = mlp
    call("opp_kpt_encode", ptr(kpts), ptr(stats), ptr(desc), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
         ptr(w3), ptr(b3), ptr(w4), ptr(b4), ptr(tok), B, N, int(split), stream())


def linear_act(a0, a1, w, out, rows, act, act_cols, split, out_split=None, batches=1, a0_shared=False,
               count=None, rows_per_count=1, row_mask=None):
    """a_i fp16 [rows, planes*k_i]; w fp16 [n, planes*(k0+k1)]; out fp16 [rows, planes*n]
    (out_split=False with split=True: split operands, single-plane output [rows, n]).
    batches > 1: rows is per batch; a0_shared: a0 is [1, rows, ..] shared by every batch.
    count (int32 device tensor): `rows` is the capacity, the kernel uses count * rows_per_count rows.
    row_mask (uint8 [batches*rows]): rows with mask 0 are written as zeros."""
    _chk(row_mask, torch.uint8, "row_mask")
    _chk(a0, torch.float16, "a0")
    _chk(a1, torch.float16, "a1")
    _chk(w, torch.float16, "w")
    planes = 2 if split else 1
    k0 = a0.shape[-1] // planes
    k1 = a1.shape[-1] // planes if a1 is not None else 0
    if split and out_split is False:
        call("opp_linear_act_f16_out1", ptr(a0), k0, ptr(a1), k1, ptr(w), ptr(out), rows, w.shape[0], act,
             act_cols, ptr(row_mask), stream())
        return out
    if count is not None:
        call("opp_linear_act_f16_dyn", ptr(a0), k0, ptr(a1), k1, ptr(w), ptr(out), rows, ptr(count),
             rows_per_count, w.shape[0], act, act_cols, int(split), stream())
        return out
    if batches > 1 or a0_shared or row_mask is not None:
        call("opp_linear_act_f16_b", ptr(a0), k0, int(a0_shared), ptr(a1), k1, ptr(w), ptr(out), batches, rows,
             w.shape[0], act, act_cols, int(split), ptr(row_mask), stream())
        return out
    call("opp_linear_act_f16", ptr(a0), k0, ptr(a1), k1, ptr(w), ptr(out), rows, w.shape[0], act,
         act_cols, int(split), stream())
    return out


def linear_q(x16, wq, ksum, out, batches, rows, v_len, split, eps=1e-6, x_shared=False, row_mask=None):
    _chk(row_mask, torch.uint8, "row_mask")
    call("opp_linear_q_f16", ptr(x16), ptr(wq), ptr(ksum), ptr(out), batches, rows, wq.shape[0],
         float(v_len), float(eps), int(split), int(x_shared), ptr(row_mask), stream())
    return out


def linear_ln(a0, a1, w, w_batched, gamma, beta, batches, rows, split, resid=None, out16=None,
              out32=None, eps=1e-5, resid_shared=False, count=None, rows_per_count=1):
    _chk(a0, torch.float16, "a0")
    _chk(w, torch.float16, "w")
    _chk(resid, torch.float16, "resid")
    planes = 2 if split else 1
    k0 = a0.shape[-1] // planes
    k1 = a1.shape[-1] // planes if a1 is not None else 0
    n = w.shape[-2]
    if count is not None:
        call("opp_linear_ln_dyn", ptr(a0), k0, ptr(a1), k1, ptr(w), ptr(gamma), ptr(beta), float(eps),
             ptr(resid), ptr(out16), ptr(out32), rows, ptr(count), rows_per_count, n, int(split), stream())
        return
    call("opp_linear_ln", ptr(a0), k0, ptr(a1), k1, ptr(w), int(w_batched), ptr(gamma), ptr(beta),
         float(eps), ptr(resid), int(resid_shared), ptr(out16), ptr(out32), batches, rows, n, int(split),
         stream())


def full_attention(q16, kv16, out, batch, l, s, heads, head_dim, split):
    """softmax(Q K^T / sqrt(D)) V per head (attention: "full"); q16 [B*l, planes*H*D], kv16 [B*s, planes*2*H*D]."""
    _chk(q16, torch.float16, "q")
    _chk(kv16, torch.float16, "kv")
    call("opp_full_attention", ptr(q16), ptr(kv16), ptr(out), batch, l, s, heads, head_dim, int(split), stream())
    return out


def kv_chunks(s, batches=None):
    """Chunks per batch element of the K'V partial states (size of the `part` buffer)."""
    if batches is None:
        return _lib.load().opp_kv_chunks(s)
    return _lib.load().opp_kv_chunks_b(s, batches)


def kv_state(kv16, part, merge_w, mt, ksum, batches, s, d, v_len, split, kv_split=None):
    """kv_split: plane mode of the kv16 rows when it differs from the mode of the mt output."""
    kv_split = split if kv_split is None else kv_split
    call("opp_kv_partial", ptr(kv16), ptr(part), batches, s, d, int(kv_split), stream())
    call("opp_kv_finalize", ptr(part), ptr(merge_w), ptr(mt), ptr(ksum), batches, kv_chunks(s, batches), d,
         float(v_len), int(split), stream())


def sim_tiles(cols):
    return _lib.load().opp_sim_tiles(cols)


def sim_lse(a, b, batches, rows, cols, k, scale, part_m, part_s, lse, split):
    tiles = sim_tiles(cols)
    call("opp_sim_lse", ptr(a), ptr(b), ptr(part_m), ptr(part_s), batches, rows, cols, k,
         float(scale), int(split), stream())
    call("opp_lse_finalize", ptr(part_m), ptr(part_s), ptr(lse), batches * rows, tiles, stream())


def sim_conf(a, b, lse_own, lse_other, own_is_pt, conf, batches, rows, cols, k, scale, part_val,
             part_idx, best_val, best_idx, split):
    tiles = sim_tiles(cols)
    call("opp_sim_conf", ptr(a), ptr(b), ptr(lse_own), ptr(lse_other), int(own_is_pt), ptr(conf),
         ptr(part_val), ptr(part_idx), batches, rows, cols, k, float(scale), int(split), stream())
    call("opp_best_finalize", ptr(part_val), ptr(part_idx), ptr(best_val), ptr(best_idx),
         batches * rows, tiles, stream())


def sim_lse_cols(a, b, batches, rows, cols, k, scale, part_m, part_s, lse_rows, col_m, col_s, lse_cols,
                 split, col_mask=None, side_stream=None):
    """lse over columns for every row (as sim_lse) AND lse over rows for every column, one GEMM pass.
    col_mask uint8 [batches, cols]: masked columns (0) get sim - 1e9 and lse_cols = +inf (conf = 0).
    side_stream (latency mode): the two independent finalisers run side by side."""
    _chk(col_mask, torch.uint8, "col_mask")
    tiles = sim_tiles(cols)
    groups = (rows + 31) // 32
    call("opp_sim_lse_cols", ptr(a), ptr(b), ptr(part_m), ptr(part_s), ptr(col_m), ptr(col_s), batches,
         rows, cols, k, float(scale), int(split), ptr(col_mask), stream())
    if side_stream is not None:
        cur = torch.cuda.current_stream()
        side_stream.wait_stream(cur)
        with torch.cuda.stream(side_stream):
            call("opp_lse_finalize", ptr(part_m), ptr(part_s), ptr(lse_rows), batches * rows, tiles, stream())
    else:
        call("opp_lse_finalize", ptr(part_m), ptr(part_s), ptr(lse_rows), batches * rows, tiles, stream())
    call("opp_lse_col_finalize", ptr(col_m), ptr(col_s), ptr(lse_cols), batches, groups, cols, ptr(col_mask),
         stream())
    if side_stream is not None:
        cur.wait_stream(side_stream)


def sim_conf_colmax(a, b, lse_own, lse_other, conf, batches, rows, cols, k, scale, part_val, part_idx,
                    best_val, best_idx, colmax, split):
    """conf pass over rows = 3D points that also leaves max_l conf[b, l, s] (float bits) in colmax."""
    tiles = sim_tiles(cols)
    call("opp_sim_conf_colmax", ptr(a), ptr(b), ptr(lse_own), ptr(lse_other), ptr(conf), ptr(part_val),
         ptr(part_idx), ptr(colmax), batches, rows, cols, k, float(scale), int(split), stream())
    call("opp_best_finalize", ptr(part_val), ptr(part_idx), ptr(best_val), ptr(best_idx),
         batches * rows, tiles, stream())


def match_select_colmax(pt_val, pt_idx, colmax, kpts, img_scale, batch, l, hc, wc, thr, border, cell,
                        scratch, b_ids, i_ids, j_ids, mconf, mkpts3d, mkpts_c, count, bank_shared=False):
    call("opp_match_select_colmax", ptr(pt_val), ptr(pt_idx), ptr(colmax), ptr(kpts), ptr(img_scale),
         batch, l, hc, wc, float(thr), int(border), float(cell), ptr(scratch), ptr(b_ids),
         ptr(i_ids), ptr(j_ids), ptr(mconf), ptr(mkpts3d), ptr(mkpts_c), ptr(count), int(bank_shared), stream())


def match_select(pt_val, pt_idx, px_idx, kpts, img_scale, batch, l, hc, wc, thr, border, cell,
                 scratch, b_ids, i_ids, j_ids, mconf, mkpts3d, mkpts_c, count, bank_shared=False):
    call("opp_match_select", ptr(pt_val), ptr(pt_idx), ptr(px_idx), ptr(kpts), ptr(img_scale),
         batch, l, hc, wc, float(thr), int(border), float(cell), ptr(scratch), ptr(b_ids),
         ptr(i_ids), ptr(j_ids), ptr(mconf), ptr(mkpts3d), ptr(mkpts_c), ptr(count), int(bank_shared), stream())


def fine_gather(fine, desc3d, b_ids, i_ids, j_ids, x32, x16, m, hf, wf, wc, stride, n, split,
                bank_shared=False, count=None, windows=False):
    """windows: `fine` is the compact [m, 5, 8, planes*128] window tensor of conv_win."""
    _chk(desc3d, torch.float32, "descriptors3d_db")
    call("opp_fine_gather", ptr(fine), ptr(desc3d), ptr(b_ids), ptr(i_ids), ptr(j_ids), ptr(x32),
         ptr(x16), m, hf, wf, wc, stride, n, int(split), int(bank_shared),
         (conv_win_pitch(5) if windows is True else int(windows)), ptr(count), stream())


def conv_win_pitch(win):
    """Row pitch of conv_win's compact output windows ([m, win, pitch, planes*Cout_pad])."""
    return _lib.load().opp_conv_win_pitch(win)


def conv_win(x, w, bias, out, win, split, m, act=0, slope=0.01, b_ids=None, j_ids=None, wc=0, stride=4,
             org=0, count=None):
    """3x3 convolution on per-match windows (opp_conv_win).  With j_ids: x is the dense NHWC map
    [B, H, W, planes*Cin_pad]; without: the compact output [m, win+2, 8, planes*Cin_pad] of the
    previous call.  out: [m, win, 8, planes*Cout_pad]."""
    _chk(x, torch.float16, "x")
    _chk(w, torch.float16, "w")
    planes = 2 if split else 1
    if j_ids is not None:
        B, H, W, C = x.shape
    else:
        B, H, W, C = 1, 0, 0, x.shape[-1]
    call("opp_conv_win", ptr(x), ptr(w), ptr(bias), ptr(out), ptr(b_ids), ptr(j_ids), m, ptr(count), B, H, W,
         C // planes, w.shape[0], win, wc, stride, org, act, float(slope), int(split), stream())
    return out


def fine_attention(qkv, msg, m, cross, split, eps=1e-6, count=None):
    call("opp_fine_attention", ptr(qkv), ptr(msg), m, int(cross), float(eps), int(split), ptr(count), stream())


def fine_match(x32, mkpts_c, b_ids, img_scale, expec_f, mkpts_f, m, fine_scale, count=None):
    call("opp_fine_match", ptr(x32), ptr(mkpts_c), ptr(b_ids), ptr(img_scale), ptr(expec_f),
         ptr(mkpts_f), m, float(fine_scale), ptr(count), stream())


# ---------------------------------------------------------------------------------------------
# LoFTR 2D-2D matcher
# ---------------------------------------------------------------------------------------------
def match_select_2d(pt_val, pt_idx, colmax, scale0, scale1, batch, h0, w0, h1, w1, thr, border, cell, scratch,
                    b_ids, i_ids, j_ids, mconf, mkpts0_c, mkpts1_c, count):
    call("opp_match_select_2d", ptr(pt_val), ptr(pt_idx), ptr(colmax), ptr(scale0), ptr(scale1), batch, h0, w0, h1,
         w1, float(thr), int(border), float(cell), ptr(scratch), ptr(b_ids), ptr(i_ids), ptr(j_ids), ptr(mconf),
         ptr(mkpts0_c), ptr(mkpts1_c), ptr(count), stream())


def fine_gather_2d(fine0, fine1, b_ids, i_ids, j_ids, x16, m, hf0, wf0, wc0, hf1, wf1, wc1, stride, window, split):
    _chk(fine0, torch.float16, "fine0")
    _chk(fine1, torch.float16, "fine1")
    call("opp_fine_gather_2d", ptr(fine0), ptr(fine1), ptr(b_ids), ptr(i_ids), ptr(j_ids), ptr(x16), m, hf0, wf0,
         wc0, hf1, wf1, wc1, stride, window, int(split), stream())


def seq_attention(q, kv, out, groups, l, s, split, eps=1e-6):
    _chk(q, torch.float16, "q")
    _chk(kv, torch.float16, "kv")
    call("opp_seq_attention", ptr(q), ptr(kv), ptr(out), groups, l, s, float(eps), int(split), stream())


def fine_match_2d(x32, mkpts1_c, b_ids, scale1, expec_f, mkpts1_f, m, window, fine_scale):
    call("opp_fine_match_2d", ptr(x32), ptr(mkpts1_c), ptr(b_ids), ptr(scale1), ptr(expec_f), ptr(mkpts1_f), m,
         window, float(fine_scale), stream())
