"""CPU tests of the host side: the C-ABI library loads and exports every symbol declared in
include/opp_b200.h, the drop-in module has the reference's state-dict layout, and the product
fails loudly (no fallback) without a GPU."""
import ctypes
import os
import pickle
import re

import pytest
import torch

from oracle import oracle, workload
from onepose_plus_plus_b200 import OnePosePlus_model, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "opp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(opp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/opp_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib.PLAIN)
    assert bound == set(syms), f"ctypes binding and header disagree: {bound ^ set(syms)}"


def test_version_and_tiles_without_gpu():
    lib = _lib.load()
    assert lib.opp_version() >= 100
    g = lib.opp_sim_tiles(100)   # partial slots per column tile = epilogue warp groups (1 or 2)
    assert g in (1, 2) and lib.opp_sim_tiles(4096) == 16 * g and lib.opp_sim_tiles(5000) == 20 * g
    assert lib.opp_kv_chunks(4096) * 128 >= 4096


def test_runtime_options_roundtrip():
    """opp_set_option / opp_get_option (include/opp_b200.h): known names toggle, unknown names fail."""
    for name in ("kv_mma", "conv1_staged", "upsample_rows", "conv1_px4", "fine_attn_vec"):
        before = _lib.get_option(name)
        assert before in (0, 1)
        _lib.set_option(name, 1 - before)
        assert _lib.get_option(name) == 1 - before
        _lib.set_option(name, before)
    assert _lib.get_option("kv_mma") == 1 and _lib.get_option("conv1_staged") == 1   # validated defaults
    for name in ("upsample_rows", "conv1_px4", "fine_attn_vec"):                        # not validated yet
        assert _lib.get_option(name) == 0
    with pytest.raises(ValueError):
        _lib.set_option("no_such_option", 1)
    assert _lib.get_option("no_such_option") == -1


def test_state_dict_is_the_reference_layout():
    m = OnePosePlus_model(oracle.DEFAULT_CONFIG)
    sd = workload.synthetic_state_dict(0)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 10_226_480
    assert "dense_pos_encoding.pe" not in m.state_dict()  # non-persistent, position_encoding.py:35
    m2 = pickle.loads(pickle.dumps(m))  # Ray ships the module object
    assert torch.equal(m2.state_dict()["backbone.conv1.weight"], sd["backbone.conv1.weight"])


def test_position_encoding_matches_oracle():
    m = OnePosePlus_model(oracle.DEFAULT_CONFIG)
    pe = oracle.position_encoding_sine(256, 24, 40)
    assert torch.allclose(m.dense_pos_encoding.pe[0, :, :24, :40], pe, atol=1e-6)


def test_no_cpu_fallback_and_config_errors():
    m = OnePosePlus_model(oracle.DEFAULT_CONFIG).eval()
    data = workload.random_workload(64, 64, 50)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(data)
    m.train()
    with pytest.raises(NotImplementedError):
        m(data)
    import copy
    bad = copy.deepcopy(oracle.DEFAULT_CONFIG)
    bad["loftr_backbone"]["type"] = "VGG"
    with pytest.raises(ValueError):
        OnePosePlus_model(bad)
    bad = copy.deepcopy(oracle.DEFAULT_CONFIG)
    bad["coarse_matching"]["type"] = "sinkhorn"
    with pytest.raises(NotImplementedError):
        OnePosePlus_model(bad)
    bad = copy.deepcopy(oracle.DEFAULT_CONFIG)
    bad["keypoints_encoding"]["type"] = "other"
    with pytest.raises(NotImplementedError):
        OnePosePlus_model(bad)


def test_bn_folding_and_planes_roundtrip():
    from onepose_plus_plus_b200 import ops
    x = torch.randn(7, 64) * 3
    for split in (0, 1):
        back = ops.from_planes(ops.to_planes(x, split), split)
        tol = 2e-6 if split else 2e-3
        assert torch.allclose(back, x, rtol=tol, atol=tol)
    m = OnePosePlus_model(oracle.DEFAULT_CONFIG)
    m.load_state_dict(workload.synthetic_state_dict(0))
    P = m._prepare(torch.device("cpu"))
    sd = m.state_dict()
    w, b = P["layer2.0.conv1"]
    assert w.shape == (208, 2 * 9 * 128) and b.shape == (208,)
    g = sd["backbone.layer2.0.bn1.weight"] / torch.sqrt(sd["backbone.layer2.0.bn1.running_var"] + 1e-5)
    ref = (sd["backbone.layer2.0.conv1.weight"] * g[:, None, None, None]).permute(0, 2, 3, 1).reshape(196, -1)
    got = ops.from_planes(w, 1)[:196]
    assert torch.allclose(got, ref, atol=1e-6)
    assert got.shape[1] == 9 * 128 and ops.from_planes(w, 1)[196:].abs().max() == 0


def test_gpu_validated_kernels_are_unchanged():
    """profiles/r1_validated_sass.txt fingerprints (sha256 of the SASS) the kernels that passed
    `pytest -m gpu`, smoke() and the bench on a B200.  An edit that changes one of them must be
    re-validated on a GPU and the file rewritten (`python scripts/sass_hash.py write ...`); new
    kernels behind off-by-default switches do not count."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_hash.py"), "check",
                        os.path.join(ROOT, "profiles", "r1_validated_sass.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]


def test_coarse_matching_host_flow(monkeypatch):
    """Stage sequencing of OnePosePlus_model._coarse_matching with the kernels stubbed out: the
    default flow is lse x2 -> conf x2 -> match_select; the one-pass switches replace exactly the
    passes they claim to, and every call binds against the real wrapper's signature."""
    import inspect
    from onepose_plus_plus_b200 import ops
    calls = []

    def stub(name):
        sig = inspect.signature(getattr(ops, name))

        def f(*a, **k):
            sig.bind(*a, **k)
            calls.append(name)
            if name.startswith("match_select"):
                a[-1].zero_()   # match count
        return f

    for n in ("sim_lse", "sim_conf", "match_select", "sim_lse_cols", "sim_conf_colmax", "match_select_colmax"):
        monkeypatch.setattr(ops, n, stub(n))
    monkeypatch.setattr(ops, "sim_tiles", lambda c: 2 * ((c + 255) // 256))
    m = OnePosePlus_model(oracle.DEFAULT_CONFIG).eval()
    assert not m.coarse_colmax and not m.coarse_lse_cols   # unvalidated paths are opt-in
    B, N, hc, wc = 2, 300, 12, 16
    q2 = torch.zeros(B, hc * wc, 512, dtype=torch.half)
    d3 = torch.zeros(B, N, 512, dtype=torch.half)
    expect = {(False, False): ["sim_lse", "sim_lse", "sim_conf", "sim_conf", "match_select"],
              (True, False): ["sim_lse", "sim_lse", "sim_conf_colmax", "match_select_colmax"],
              (False, True): ["sim_lse_cols", "sim_conf", "sim_conf", "match_select"],
              (True, True): ["sim_lse_cols", "sim_conf_colmax", "match_select_colmax"]}
    for flags, want in expect.items():
        m.coarse_colmax, m.coarse_lse_cols = flags
        calls.clear()
        data = {"keypoints3d": torch.zeros(B, N, 3), "query_image_scale": torch.ones(B, 2),
                "q_hw_i": torch.Size((96, 128))}
        M, _ = m._coarse_matching(q2, d3, data, B, N, hc, wc)
        assert M == 0 and calls == want
        assert data["conf_matrix"].shape == (B, N, hc * wc) and data["b_ids"].numel() == 0


def test_full_forward_host_flow(monkeypatch):
    """Every stage of the forward with all kernels stubbed (CPU tensors): checks the host-side
    sequencing, buffer shapes and wrapper signatures of backbone -> kpt encode -> coarse transformer
    -> coarse matching -> fine stage, and the launch count per forward the bench reports."""
    import inspect
    from onepose_plus_plus_b200 import ops
    calls = []
    returns_out = {"conv1_7x7": 3, "conv2d_nhwc": 3, "upsample2x_add": 2, "linear_act": 3, "linear_q": 3}

    def stub(name):
        sig = inspect.signature(getattr(ops, name))

        def f(*a, **k):
            bound = sig.bind(*a, **k)
            calls.append(name)
            if name.startswith("match_select"):
                bound.arguments["count"].fill_(5)
            if name in returns_out:
                return a[returns_out[name]]
        return f

    names = [n for n, fn in inspect.getmembers(ops, inspect.isfunction)
             if n not in ("to_planes", "from_planes", "_chk", "kv_chunks", "sim_tiles")]
    for n in names:
        monkeypatch.setattr(ops, n, stub(n))
    monkeypatch.setattr(ops, "sim_tiles", lambda c: 2 * ((c + 255) // 256))
    monkeypatch.setattr(ops, "kv_chunks", lambda s_: (s_ + 127) // 128)
    m = OnePosePlus_model(oracle.DEFAULT_CONFIG).eval()
    m.load_state_dict(workload.synthetic_state_dict(0))
    dev = torch.device("cpu")
    m._plan = m._prepare(dev)
    B, H, W, N = 2, 64, 96, 200
    img = torch.rand(B, 1, H, W)
    q2, fine_map, (hc, wc) = m._backbone(img)
    assert (hc, wc) == (H // 8, W // 8) and q2.shape == (B, hc * wc, 512)
    assert fine_map.shape == (B, H // 2, W // 2, 256)
    assert calls.count("conv2d_nhwc") == 21 and calls.count("upsample2x_add") == 2 and calls.count("conv1_7x7") == 1
    calls.clear()
    d3 = torch.zeros(B, N, 512, dtype=torch.half)
    o2, o3 = m._coarse_transformer(q2, d3, B, hc * wc, N)
    assert o2.shape == q2.shape and o3.shape == d3.shape
    # 6 layers x 2 sequences x (kv GEMM, kv_state, q GEMM, Mt+LN, mlp0, mlp2+LN)
    assert calls.count("linear_act") == 24 and calls.count("linear_ln") == 24
    assert calls.count("linear_q") == 12 and calls.count("kv_state") == 12
    m.kv_single_plane = True    # opt-in: K'/V rows as one fp16 plane (same launches, shorter rows)
    calls.clear()
    m._coarse_transformer(q2, d3, B, hc * wc, N)
    assert calls.count("linear_act") == 24 and calls.count("kv_state") == 12
    assert m._buf("c2_kv16", (B * hc * wc, 512), torch.float16, dev).shape[1] == 512
    m.kv_single_plane = False
    calls.clear()
    data = {"keypoints3d": torch.zeros(B, N, 3), "query_image_scale": torch.ones(B, 2),
            "q_hw_i": img.shape[2:], "q_hw_c": torch.Size((hc, wc)), "q_hw_f": torch.Size(fine_map.shape[1:3]),
            "descriptors3d_db": torch.zeros(B, 128, N)}
    M, img_scale = m._coarse_matching(o2, o3, data, B, N, hc, wc)
    assert M == 5 and data["b_ids"].shape == (5,) and data["mkpts_3d_db"].shape == (5, 3)
    calls.clear()
    m._fine(data, fine_map, M, img_scale, wc)
    assert data["expec_f"].shape == (5, 3) and data["mkpts_query_f"].shape == (5, 2) and data["W"] == 5
    assert calls == ["fine_gather"] + ["linear_act", "fine_attention", "linear_ln", "linear_act", "linear_ln"] * 2 \
        + ["fine_match"]
    # empty-match path (fine_preprocess.py:34-37, fine_matching.py:46-55)
    data["mkpts_query_c"] = torch.zeros(0, 2)
    m._fine(data, fine_map, 0, img_scale, wc)
    assert data["expec_f"].shape == (0, 3) and data["mkpts_query_f"].shape == (0, 2)
