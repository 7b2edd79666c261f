"""ctypes binding of ``libopp_b200.so`` (C ABI declared in ``include/opp_b200.h``).

There is no fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is
raised.  Tensors are passed as raw device pointers; the current torch CUDA stream is used.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OPP_B200_LIB") or os.path.join(_HERE, "libopp_b200.so")

_lib = None


def _sig(fn, argtypes):
    fn.argtypes = argtypes
    fn.restype = c_int


P, I, L, F = c_void_p, c_int, c_longlong, c_float

# name -> argtypes; must mirror include/opp_b200.h exactly (tests check every symbol resolves)
SIGNATURES = {
    "opp_conv2d_nhwc": [P, P, P, P, P, I, I, I, I, I, I, I, I, F, P, P, P, I, P],
    "opp_conv1_im2col": [P, I, P, I, I, I, I, P],
    "opp_kpt_stats": [P, P, I, I, P],
    "opp_kpt_encode": [P] * 12 + [I, I, I, P],
    "opp_linear_act_f16": [P, I, P, I, P, P, L, I, I, I, I, P],
    "opp_linear_act_f16_out1": [P, I, P, I, P, P, L, I, I, I, P, P],
    "opp_linear_act_f16_b": [P, I, I, P, I, P, P, I, L, I, I, I, I, P, P],
    "opp_linear_q_f16": [P, P, P, P, I, I, I, F, F, I, I, P, P],
    "opp_linear_ln": [P, I, P, I, P, I, P, P, F, P, I, P, P, I, L, I, I, P],
    "opp_full_attention": [P, P, P, I, I, I, I, I, I, P],
    "opp_kv_partial": [P, P, I, I, I, I, P],
    "opp_kv_finalize": [P, P, P, P, I, I, I, F, I, P],
    "opp_sim_lse": [P, P, P, P, I, I, I, I, F, I, P],
    "opp_lse_finalize": [P, P, P, L, I, P],
    "opp_sim_conf": [P, P, P, P, I, P, P, P, I, I, I, I, F, I, P],
    "opp_sim_lse_cols": [P, P, P, P, P, P, I, I, I, I, F, I, P, P],
    "opp_lse_col_finalize": [P, P, P, I, I, I, P, P],
    "opp_sim_conf_colmax": [P, P, P, P, P, P, P, P, I, I, I, I, F, I, P],
    "opp_best_finalize": [P, P, P, P, L, I, P],
    "opp_match_select": [P, P, P, P, P, I, I, I, I, F, I, F, P, P, P, P, P, P, P, P, I, P],
    "opp_match_select_colmax": [P, P, P, P, P, I, I, I, I, F, I, F, P, P, P, P, P, P, P, P, I, P],
    "opp_fine_gather": [P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P, P],
    "opp_conv_win": [P, P, P, P, P, P, I, P, I, I, I, I, I, I, I, I, I, I, F, I, P],
    "opp_fine_attention": [P, P, I, I, F, I, P, P],
    "opp_fine_match": [P, P, P, P, P, P, I, F, P, P],
    "opp_linear_act_f16_dyn": [P, I, P, I, P, P, L, P, I, I, I, I, I, P],
    "opp_linear_ln_dyn": [P, I, P, I, P, P, P, F, P, P, P, L, P, I, I, I, P],
    "opp_match_select_2d": [P, P, P, P, P, I, I, I, I, I, F, I, F, P, P, P, P, P, P, P, P, P],
    "opp_fine_gather_2d": [P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, P],
    "opp_seq_attention": [P, P, P, I, I, I, F, I, P],
    "opp_fine_match_2d": [P, P, P, P, P, P, I, I, F, P],
    "opp_pnp_ransac": [P, P, P, I, P, I, F, F, I, ctypes.c_uint, I, P, P, P, P, P],
}
PLAIN = {"opp_version": ([], c_int), "opp_num_sms": ([], c_int), "opp_sim_tiles": ([I], c_int),
         "opp_kv_chunks": ([I], c_int),
         "opp_kv_chunks_b": ([I, I], c_int),
         "opp_conv_win_pitch": ([I], c_int),
         "opp_last_error": ([], ctypes.c_char_p)}


def load():
    """Load the shared library (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU/PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        _sig(getattr(lib, name), argtypes)
    for name, (argtypes, restype) in PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# kernels launched per entry point (bench.py reports the per-step total as gpu_launches)
KERNELS_PER_CALL = {"opp_match_select": 3, "opp_match_select_colmax": 3}
LAUNCHES = 0
_PROFILE = None


def call(name, *args):
    global LAUNCHES
    lib = load()
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.opp_last_error()
        raise RuntimeError(f"{name} failed (status {rc}): {msg.decode() if msg else ''}")
    LAUNCHES += KERNELS_PER_CALL.get(name, 1)
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append((name, e0, e1))


def profile_ops(fn, out):
    """Run fn() once with CUDA events around every C-ABI call and print per-entry-point totals."""
    global _PROFILE
    _PROFILE = []
    fn()
    torch.cuda.synchronize()
    rows, _PROFILE_local = {}, _PROFILE
    _PROFILE = None
    for name, e0, e1 in _PROFILE_local:
        ms = e0.elapsed_time(e1)
        n, t = rows.get(name, (0, 0.0))
        rows[name] = (n + 1, t + ms)
    total = sum(t for _, t in rows.values())
    for name, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:24s} calls={n:4d} total={t:9.3f} ms  {100 * t / total:5.1f}%", file=out)
    print(f"{'sum':24s} {'':10s} total={total:9.3f} ms", file=out)
    return rows
