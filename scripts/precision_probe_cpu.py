"""CPU experiment behind DESIGN.md §2: inject fp16 rounding at one intermediate tensor of the
coarse transformer (oracle, fp32 elsewhere) and report how far conf_matrix / the match list move.
    python scripts/precision_probe_cpu.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, workload  # noqa: E402

sd = workload.synthetic_state_dict(0)
FLAGS = {}


def R(t):
    return t.half().float()


def attention(q, k, v, eps=1e-6):
    Q, K = F.elu(q) + 1, F.elu(k) + 1
    coarse = q.shape[-1] * q.shape[-2] == 256
    if coarse and FLAGS.get("kv"):
        K, v = R(K), R(v)
    vl = v.size(1)
    KV = torch.einsum("nshd,nshv->nhdv", K, v / vl)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(1)) + eps)
    QZ = Q * Z[..., None] * vl
    if coarse and FLAGS.get("qz"):
        QZ = R(QZ)
    return torch.einsum("nlhd,nhdv->nlhv", QZ, KV)


def encoder_layer(sd_, p, x, source, nhead):
    bs, d = x.size(0), x.size(2)
    dim, coarse = d // nhead, d == 256
    q = F.linear(x, sd_[p + "q_proj.weight"]).view(bs, -1, nhead, dim)
    k = F.linear(source, sd_[p + "k_proj.weight"]).view(bs, -1, nhead, dim)
    v = F.linear(source, sd_[p + "v_proj.weight"]).view(bs, -1, nhead, dim)
    msg = attention(q, k, v).reshape(bs, -1, d)
    msg = F.layer_norm(F.linear(msg, sd_[p + "merge.weight"]), (d,), sd_[p + "norm1.weight"], sd_[p + "norm1.bias"], 1e-5)
    if coarse and FLAGS.get("msg"):
        msg = R(msg)
    h = F.relu(F.linear(torch.cat([x, msg], 2), sd_[p + "mlp.0.weight"]))
    if coarse and FLAGS.get("h"):
        h = R(h)
    msg = F.layer_norm(F.linear(h, sd_[p + "mlp.2.weight"]), (d,), sd_[p + "norm2.weight"], sd_[p + "norm2.bias"], 1e-5)
    return x + msg


data, _ = workload.planted_workload(sd, 512, 512, 5000, 3000, batch=1)
ref = {k: v.clone() for k, v in data.items()}
oracle.forward(sd, ref)
oracle.encoder_layer = encoder_layer
for flags in ({}, {"kv": 1}, {"qz": 1}, {"msg": 1}, {"h": 1}, {"kv": 1, "qz": 1, "msg": 1, "h": 1}):
    FLAGS.clear()
    FLAGS.update(flags)
    got = {k: v.clone() for k, v in data.items()}
    oracle.forward(sd, got)
    same = all(torch.equal(ref[k], got[k]) for k in ("b_ids", "i_ids", "j_ids"))
    print(f"{str(flags):44s} M={got['b_ids'].numel()} indices_equal={same} "
          f"conf_max_change={(ref['conf_matrix'] - got['conf_matrix']).abs().max().item():.2e}")
