"""ctypes binding of libtrxlppo.so (C ABI in include/trxl_ppo.h).

PyTorch is only the owner of device memory and streams here: every wrapper passes raw
``tensor.data_ptr()`` values plus the current CUDA stream handle.  There is NO CPU fallback: if the
shared library is missing or a symbol is absent, importing callers fail with an error that says how
to build it (``python __graft_entry__.py``)."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrxlppo.so")
MAX_BRANCHES = 8
LN_MODES = {"": 0, None: 0, "pre": 1, "post": 2}
PE_MODES = {"": 0, None: 0, "relative": 1, "learned": 2}

_lib = None

vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double


class ModelConfig(C.Structure):
    _fields_ = [("embed_dim", C.c_int32), ("num_heads", C.c_int32), ("num_blocks", C.c_int32),
                ("memory_length", C.c_int32), ("hidden_size", C.c_int32), ("feat_dim", C.c_int32),
                ("layer_norm", C.c_int32), ("pos_enc", C.c_int32), ("gtrxl", C.c_int32),
                ("max_episode_steps", C.c_int32), ("num_branches", C.c_int32),
                ("branch_sizes", C.c_int32 * MAX_BRANCHES), ("conv_in_channels", C.c_int32)]


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("ndim", C.c_int32), ("shape", C.c_int64 * 4),
                ("group", C.c_int32)]


class AttnGroups(C.Structure):
    _fields_ = [("table_pe", C.c_void_p), ("n_episodes", C.c_int32), ("tiles", C.c_void_p), ("n_tiles", C.c_int32),
                ("ranges", C.c_void_p)]


CFGP = C.POINTER(ModelConfig)
GRPP = C.POINTER(AttnGroups)

# name -> (restype, argtypes); this is the full exported surface of include/trxl_ppo.h
SIGNATURES = {
    "trxl_last_error": (C.c_char_p, []),
    "trxl_abi_version": (i32, []),
    "trxl_launch_count": (i64, []),
    "trxl_tc_gemm_launches": (i64, []),
    "trxl_profile_enable": (i32, [i32]),
    "trxl_profile_read": (i32, [i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "trxl_profile_aux": (i64, [i32, i32]),
    "trxl_graph_begin": (i32, [vp]),
    "trxl_graph_end": (i32, [vp, C.POINTER(C.c_void_p)]),
    "trxl_graph_launch": (i32, [vp, vp]),
    "trxl_graph_destroy": (i32, [vp]),
    "trxl_copy_rows": (i32, [vp, vp, i64, i64, i64, i64, vp]),
    "trxl_copy_async": (i32, [vp, vp, i64, vp]),
    "trxl_layout_num_entries": (i32, [CFGP]),
    "trxl_layout_total_floats": (i64, [CFGP]),
    "trxl_layout_entry": (i32, [CFGP, i32, C.POINTER(ParamEntry)]),
    "trxl_layout_groups": (i32, [CFGP]),
    "trxl_workspace_floats": (i64, [CFGP, i32]),
    "trxl_fused_forward_supported": (i32, [CFGP]),
    "trxl_model_forward": (i32, [CFGP, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    "trxl_model_backward": (i32, [CFGP, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
    "trxl_model_forward_grouped": (i32, [CFGP, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, GRPP, vp]),
    "trxl_model_backward_grouped": (i32, [CFGP, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, GRPP, vp]),
    "trxl_grouped_attention_supported": (i32, [CFGP]),
    "trxl_table_add_pe": (i32, [vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "trxl_attention_ranges": (i32, [vp, vp, vp, vp, i32, i32, vp, vp]),
    "trxl_conv_encoder_workspace_floats": (i64, [CFGP, i32, i32, i32]),
    "trxl_conv_encoder_forward": (i32, [CFGP, vp, vp, i32, i32, i32, vp, vp, vp]),
    "trxl_conv_train_supported": (i32, [CFGP, i32, i32]),
    "trxl_conv_train_workspace_floats": (i64, [CFGP, i32, i32, i32]),
    "trxl_conv_train_pack_weights": (i32, [CFGP, vp, i32, i32, i32, vp, vp]),
    "trxl_conv_train_forward": (i32, [CFGP, vp, vp, vp, i32, i32, i32, vp, vp, i32, vp]),
    "trxl_conv_train_backward": (i32, [CFGP, vp, i32, i32, i32, vp, vp, vp]),
    "trxl_window_attention_forward": (i32, [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "trxl_window_attention_backward": (i32, [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                             vp, vp, vp, vp]),
    "trxl_linear_forward": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "trxl_linear_backward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "trxl_layernorm_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "trxl_layernorm_backward": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "trxl_gather_window": (i32, [vp, vp, vp, i64, i32, i64, i64, vp]),
    "trxl_gather_rows": (i32, [vp, vp, vp, i64, i64, vp]),
    "trxl_gae": (i32, [vp, vp, vp, vp, vp, i32, i32, f64, f64, vp]),
    "trxl_rollout_prepare": (i32, [vp, vp, vp, vp, vp, i64, vp, i64, vp, i64, i32, i32, vp]),
    "trxl_memory_scatter": (i32, [vp, vp, vp, vp, i32, i64, i64, vp]),
    "trxl_rollout_store": (i32, [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, vp, vp, i64, vp]),
    "trxl_rollout_fetch": (i32, [vp, i64, vp, vp, vp, vp, i64, vp, vp, i32, vp]),
    "trxl_host_device_pointer": (i32, [vp, C.POINTER(C.c_void_p)]),
    "trxl_sample_actions": (i32, [vp, vp, vp, C.POINTER(C.c_int32), i32, vp, i64, vp, i64, vp, i32, vp]),
    "trxl_sample_actions_notify": (i32, [vp, vp, vp, C.POINTER(C.c_int32), i32, vp, i64, vp, i64, vp, i32, vp, vp, vp]),
    "trxl_adv_stats": (i32, [vp, vp, i32, vp, vp]),
    "trxl_ppo_loss": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int32), i32, i32, f64, f64, f64, vp, vp, vp, vp, vp]),
    "trxl_clip_adamw_step": (i32, [vp, vp, vp, vp, i64, vp, i32, i32, f64, f64, f64, f64, f64, f64, i64, vp, vp, vp]),
    "trxl_comm_unique_id": (i32, [vp]),
    "trxl_comm_create": (i32, [vp, i32, i32, C.POINTER(C.c_void_p)]),
    "trxl_comm_destroy": (i32, [vp]),
    "trxl_comm_calls": (i64, [vp]),
    "trxl_comm_nccl_version": (i32, []),
    "trxl_allreduce_grads": (i32, [vp, vp, i64, vp]),
    "trxl_allreduce_f64": (i32, [vp, vp, i64, vp]),
}
COMM_ID_BYTES = 128


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load libtrxlppo.so and bind every symbol of the ABI.  Raises NativeLibraryError (never
    falls back to another implementation) when the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError("%s not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                                 "This engine has no CPU/PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError("libtrxlppo.so does not export %s (stale build?)" % name) from e
        fn.restype, fn.argtypes = res, args
    if lib.trxl_abi_version() != 1:
        raise NativeLibraryError("libtrxlppo.so ABI version mismatch")
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, _lib.trxl_last_error().decode()))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    """Device pointer of a tensor (None -> NULL).  Tensors must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeLibraryError("libtrxlppo operates on CUDA tensors only (got %s); no CPU fallback exists" % t.device)
    assert t.is_contiguous(), "non-contiguous tensor passed to native op"
    return t.data_ptr()


def branch_array(sizes):
    arr = (C.c_int32 * len(sizes))(*[int(s) for s in sizes])
    return arr


def make_config(embed_dim, num_heads, num_blocks, memory_length, hidden_size, feat_dim, layer_norm, pos_enc, gtrxl,
                max_episode_steps, branch_sizes, conv_in_channels=0):
    cfg = ModelConfig()
    cfg.embed_dim, cfg.num_heads, cfg.num_blocks, cfg.memory_length = embed_dim, num_heads, num_blocks, memory_length
    cfg.hidden_size, cfg.feat_dim = hidden_size, feat_dim
    cfg.layer_norm = LN_MODES[layer_norm] if not isinstance(layer_norm, int) else layer_norm
    cfg.pos_enc = PE_MODES[pos_enc] if not isinstance(pos_enc, int) else pos_enc
    cfg.gtrxl = int(bool(gtrxl))
    cfg.max_episode_steps = int(max_episode_steps)
    cfg.num_branches = len(branch_sizes)
    for i, s in enumerate(branch_sizes):
        cfg.branch_sizes[i] = int(s)
    cfg.conv_in_channels = int(conv_in_channels)
    return cfg


def layout(cfg):
    """[(name, offset, shape, group)], total floats, number of grad-norm groups."""
    lib = load()
    n = lib.trxl_layout_num_entries(C.byref(cfg))
    if n < 0:
        raise ValueError("invalid model config: " + lib.trxl_last_error().decode())
    out = []
    e = ParamEntry()
    for i in range(n):
        _check(lib.trxl_layout_entry(C.byref(cfg), i, C.byref(e)), "trxl_layout_entry")
        out.append((e.name.decode(), int(e.offset), tuple(int(e.shape[k]) for k in range(e.ndim)), int(e.group)))
    return out, int(lib.trxl_layout_total_floats(C.byref(cfg))), int(lib.trxl_layout_groups(C.byref(cfg)))


def launch_count():
    return int(load().trxl_launch_count())


def graph_begin(stream_handle):
    _check(load().trxl_graph_begin(stream_handle), "trxl_graph_begin")


def graph_end(stream_handle):
    out = C.c_void_p(None)
    _check(load().trxl_graph_end(stream_handle, C.byref(out)), "trxl_graph_end")
    return out.value


def graph_abort(stream_handle):
    out = C.c_void_p(None)
    load().trxl_graph_end(stream_handle, C.byref(out))
    if out.value:
        load().trxl_graph_destroy(out.value)


def graph_launch(graph_exec):
    _check(load().trxl_graph_launch(graph_exec, _stream()), "trxl_graph_launch")


def graph_launch_on(graph_exec, stream_handle):
    """Replay on an explicit stream handle (no torch stream lookup: the rollout's per-step fast path)."""
    if _lib.trxl_graph_launch(graph_exec, stream_handle) != 0:
        raise RuntimeError("trxl_graph_launch failed: %s" % _lib.trxl_last_error().decode())


def graph_destroy(graph_exec):
    load().trxl_graph_destroy(graph_exec)


def copy_rows(src_ptr, dst_ptr, rows, row_bytes, src_stride_bytes, dst_stride_bytes):
    _check(load().trxl_copy_rows(src_ptr, dst_ptr, int(rows), int(row_bytes), int(src_stride_bytes), int(dst_stride_bytes), _stream()),
           "trxl_copy_rows")


def copy_async(src_ptr, dst_ptr, nbytes):
    _check(load().trxl_copy_async(src_ptr, dst_ptr, int(nbytes), _stream()), "trxl_copy_async")


def tc_gemm_launches():
    return int(load().trxl_tc_gemm_launches())


_profiling = False


def profile_enable(on):
    global _profiling
    _check(load().trxl_profile_enable(int(bool(on))), "trxl_profile_enable")
    _profiling = bool(on)


def profiling():
    """True while the library's event timers are on (they cannot record inside a captured graph)."""
    return _profiling


def profile_read(kind, min_samples=0):
    """(total_ms, launches, samples) of the timed attention launches of `kind` (0 fwd, 1 bwd)."""
    ms, launches, samples = C.c_double(0), C.c_int64(0), C.c_int64(0)
    _check(load().trxl_profile_read(int(kind), int(min_samples), C.byref(ms), C.byref(launches), C.byref(samples)),
           "trxl_profile_read")
    return ms.value, launches.value, samples.value


def profile_aux(kind, min_samples=0):
    return int(load().trxl_profile_aux(int(kind), int(min_samples)))


def fused_forward_supported(cfg):
    return bool(load().trxl_fused_forward_supported(C.byref(cfg)))


def workspace_floats(cfg, n):
    v = load().trxl_workspace_floats(C.byref(cfg), int(n))
    if v < 0:
        raise ValueError("invalid config: " + _lib.trxl_last_error().decode())
    return int(v)


def attn_groups(table_pe, n_episodes, tiles, n_tiles, ranges):
    """trxl_attn_groups for the episode-grouped tensor-core attention (device tensors; the struct only holds pointers, keep the
    tensors alive while it is in use)."""
    g = AttnGroups()
    g.table_pe, g.n_episodes, g.tiles, g.n_tiles, g.ranges = _p(table_pe), int(n_episodes), _p(tiles), int(n_tiles), _p(ranges)
    return g


def grouped_attention_supported(cfg):
    return bool(load().trxl_grouped_attention_supported(C.byref(cfg)))


def table_add_pe(table, pe_table, out, n_episodes, layer_norm=False):
    _, m, b, d = table.shape
    _check(load().trxl_table_add_pe(_p(table), _p(pe_table), _p(out), int(n_episodes), m, b, d, int(bool(layer_norm)), _stream()),
           "trxl_table_add_pe")


def attention_ranges(mask, win_index, ep_index, sample_index, n, L, ranges):
    _check(load().trxl_attention_ranges(_p(mask), _p(win_index), _p(ep_index), _p(sample_index), int(n), int(L), _p(ranges),
                                        _stream()), "trxl_attention_ranges")


def model_forward(cfg, params, feat, table, slots, ep_index, win_index, mask, pe_index, sample_index, pe_table, n, ws,
                  logits, value, out_mem, groups=None):
    lib = load()
    if groups is not None:
        _check(lib.trxl_model_forward_grouped(C.byref(cfg), _p(params), _p(feat), _p(table), int(slots), _p(ep_index), _p(win_index),
                                              _p(mask), _p(pe_index), _p(sample_index), _p(pe_table), int(n), _p(ws), _p(logits),
                                              _p(value), _p(out_mem), C.byref(groups), _stream()), "trxl_model_forward_grouped")
        return
    _check(lib.trxl_model_forward(C.byref(cfg), _p(params), _p(feat), _p(table), int(slots), _p(ep_index), _p(win_index),
                                  _p(mask), _p(pe_index), _p(sample_index), _p(pe_table), int(n), _p(ws), _p(logits),
                                  _p(value), _p(out_mem), _stream()), "trxl_model_forward")


def model_backward(cfg, params, grads, feat, table, slots, ep_index, win_index, mask, pe_index, sample_index, pe_table, n, ws,
                   out_mem, dlogits, dvalue, dfeat, groups=None):
    lib = load()
    if groups is not None:
        _check(lib.trxl_model_backward_grouped(C.byref(cfg), _p(params), _p(grads), _p(feat), _p(table), int(slots), _p(ep_index),
                                               _p(win_index), _p(mask), _p(pe_index), _p(sample_index), _p(pe_table), int(n), _p(ws),
                                               _p(out_mem), _p(dlogits), _p(dvalue), _p(dfeat), C.byref(groups), _stream()),
               "trxl_model_backward_grouped")
        return
    _check(lib.trxl_model_backward(C.byref(cfg), _p(params), _p(grads), _p(feat), _p(table), int(slots), _p(ep_index),
                                   _p(win_index), _p(mask), _p(pe_index), _p(sample_index), _p(pe_table), int(n), _p(ws),
                                   _p(out_mem), _p(dlogits), _p(dvalue), _p(dfeat), _stream()), "trxl_model_backward")


def conv_encoder_workspace_floats(cfg, n, h, w):
    v = load().trxl_conv_encoder_workspace_floats(C.byref(cfg), int(n), int(h), int(w))
    if v < 0:
        raise ValueError("no conv encoder for this config / observation too small")
    return int(v)


def conv_encoder_forward(cfg, params, obs, ws, feat):
    n, _, h, w = obs.shape
    _check(load().trxl_conv_encoder_forward(C.byref(cfg), _p(params), _p(obs), n, h, w, _p(ws), _p(feat), _stream()),
           "trxl_conv_encoder_forward")


def conv_train_supported(cfg, h, w):
    return bool(load().trxl_conv_train_supported(C.byref(cfg), int(h), int(w)))


def conv_train_workspace_floats(cfg, n, h, w):
    v = load().trxl_conv_train_workspace_floats(C.byref(cfg), int(n), int(h), int(w))
    if v < 0:
        raise ValueError("the tensor-core encoder does not cover this observation shape")
    return int(v)


def conv_train_pack_weights(cfg, params, n, h, w, ws):
    _check(load().trxl_conv_train_pack_weights(C.byref(cfg), _p(params), int(n), int(h), int(w), _p(ws), _stream()),
           "trxl_conv_train_pack_weights")


def conv_train_forward(cfg, params, obs, sample_index, n, ws, feat, repack=True):
    h, w = obs.shape[-2:]
    _check(load().trxl_conv_train_forward(C.byref(cfg), _p(params), _p(obs), _p(sample_index), int(n), int(h), int(w), _p(ws),
                                          _p(feat), 1 if repack else 0, _stream()), "trxl_conv_train_forward")


def conv_train_backward(cfg, grads, n, h, w, ws, dfeat):
    _check(load().trxl_conv_train_backward(C.byref(cfg), _p(grads), int(n), int(h), int(w), _p(ws), _p(dfeat), _stream()),
           "trxl_conv_train_backward")


def window_attention_forward(table, slots, num_blocks, block, ep_index, win_index, mask, pe_index, sample_index, pe_table, qk,
                             qkb, ln, n, L, D, H, probs, ctx):
    lib = load()
    _check(lib.trxl_window_attention_forward(_p(table), int(slots), num_blocks, block, _p(ep_index), _p(win_index), _p(mask),
                                             _p(pe_index), _p(sample_index), _p(pe_table), _p(qk), _p(qkb), int(ln), n, L, D, H,
                                             _p(probs), _p(ctx), _stream()), "trxl_window_attention_forward")


def window_attention_backward(table, slots, num_blocks, block, ep_index, win_index, mask, pe_index, sample_index, pe_table, qk,
                              probs, ctx, dctx, ln, n, L, D, H, dqk, dqkb, dpe):
    lib = load()
    _check(lib.trxl_window_attention_backward(_p(table), int(slots), num_blocks, block, _p(ep_index), _p(win_index), _p(mask),
                                              _p(pe_index), _p(sample_index), _p(pe_table), _p(qk), _p(probs), _p(ctx),
                                              _p(dctx), int(ln), n, L, D, H, _p(dqk), _p(dqkb), _p(dpe), _stream()),
           "trxl_window_attention_backward")


def linear_forward(x, w, bias, y, relu=False):
    m, k = x.shape
    n = w.shape[0]
    _check(load().trxl_linear_forward(_p(x), _p(w), _p(bias), _p(y), m, n, k, int(relu), _stream()), "trxl_linear_forward")


def linear_backward(dy, x, w, dx, dw, db, scratch):
    m, n = dy.shape
    k = w.shape[1] if w is not None else x.shape[1]
    _check(load().trxl_linear_backward(_p(dy), _p(x), _p(w), _p(dx), _p(dw), _p(db), m, n, k, _p(scratch), _stream()),
           "trxl_linear_backward")


def layernorm_forward(x, gamma, beta, y, mean, rstd):
    rows, d = x.shape
    _check(load().trxl_layernorm_forward(_p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, d, _stream()),
           "trxl_layernorm_forward")


def layernorm_backward(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, scratch):
    rows, d = x.shape
    _check(load().trxl_layernorm_backward(_p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dx), _p(dgamma), _p(dbeta),
                                          _p(scratch), rows, d, _stream()), "trxl_layernorm_backward")


def gather_window(src, index, out):
    n, slots = src.shape[0], src.shape[1]
    inner = src[0, 0].numel()
    _check(load().trxl_gather_window(_p(src), _p(index), _p(out), n, index.shape[1], slots, inner, _stream()), "trxl_gather_window")


def gather_rows(src, index, dst):
    rows = index.shape[0]
    row_floats = src[0].numel()
    _check(load().trxl_gather_rows(_p(src), _p(index), _p(dst), rows, row_floats, _stream()), "trxl_gather_rows")


def gae(rewards, dones, values, last_value, adv, gamma, lamda):
    w, t = values.shape
    _check(load().trxl_gae(_p(rewards), _p(dones), _p(values), _p(last_value), _p(adv), w, t, float(gamma), float(lamda),
                           _stream()), "trxl_gae")


def rollout_prepare(step, ep, mask_table, index_table, mask_out_ptr, mask_stride, idx_out_ptr, idx_stride, ep_out_ptr,
                    ep_stride, w, L):
    _check(load().trxl_rollout_prepare(_p(step), _p(ep), _p(mask_table), _p(index_table), mask_out_ptr, mask_stride,
                                       idx_out_ptr, idx_stride, ep_out_ptr, ep_stride, w, L, _stream()), "trxl_rollout_prepare")


def rollout_fetch(obs_src_ptr, obs_floats, step_src_ptr, ep_src_ptr, obs_dev, obs_store_ptr, store_stride_floats, step_dev, ep_dev, n):
    _check(load().trxl_rollout_fetch(obs_src_ptr, int(obs_floats), step_src_ptr, ep_src_ptr, _p(obs_dev), obs_store_ptr,
                                     int(store_stride_floats), _p(step_dev), _p(ep_dev), int(n), _stream()), "trxl_rollout_fetch")


def host_device_pointer(host_ptr):
    """Device-side address of pinned / cudaHostRegister'ed host memory (raises if the buffer is not mapped)."""
    out = C.c_void_p(None)
    _check(load().trxl_host_device_pointer(host_ptr, C.byref(out)), "trxl_host_device_pointer")
    return out.value


def memory_scatter(table, ep, step, new_mem, slots, inner):
    _check(load().trxl_memory_scatter(_p(table), _p(ep), _p(step), _p(new_mem), ep.shape[0], int(slots), int(inner), _stream()),
           "trxl_memory_scatter")


def rollout_store(table, table_pe, pe_table, ep, step, new_mem, slots, blocks, dim, value=None, value_dst=0, value_stride=0):
    """End-of-step stores in one kernel: memory rows into `table` (and, with their positional rows, into `table_pe` if given) and
    `value` into the rollout buffer (`value_dst`: raw device address of the first worker's slot, stride in floats)."""
    _check(load().trxl_rollout_store(_p(table), _p(table_pe), _p(pe_table), _p(ep), _p(step), _p(new_mem), ep.shape[0], int(slots),
                                     int(blocks), int(dim), _p(value), int(value_dst) or None, int(value_stride),
                                     _stream()), "trxl_rollout_store")


def sample_actions(logits, u, branch_sizes, act_ptr, act_stride, logp_ptr, logp_stride, act_compact, w, forced=None, notify=None):
    """``act_compact``: (w, nb) int64 device tensor, or the raw device-side address of a mapped host buffer.
    ``notify`` = (device counter tensor, device-side address of a pinned host int64 or None): the compact action words are
    written as ``seq << 32 | action`` for host polling; the optional flag is published behind system-wide fences."""
    compact = act_compact if isinstance(act_compact, int) else _p(act_compact)
    if notify is not None:
        _check(load().trxl_sample_actions_notify(_p(logits), _p(u), _p(forced), branch_array(branch_sizes), len(branch_sizes), act_ptr,
                                                 act_stride, logp_ptr, logp_stride, compact, w, _p(notify[0]), notify[1], _stream()),
               "trxl_sample_actions_notify")
        return
    _check(load().trxl_sample_actions(_p(logits), _p(u), _p(forced), branch_array(branch_sizes), len(branch_sizes), act_ptr, act_stride,
                                      logp_ptr, logp_stride, compact, w, _stream()), "trxl_sample_actions")


def adv_stats(adv, sample_index, n, out3):
    _check(load().trxl_adv_stats(_p(adv), _p(sample_index), n, _p(out3), _stream()), "trxl_adv_stats")


def ppo_loss(logits, value, actions, old_logp, old_values, adv, sample_index, adv_stats3, branch_sizes, n, clip, beta, vf_coef,
             dlogits, dvalue, stats6, scratch):
    _check(load().trxl_ppo_loss(_p(logits), _p(value), _p(actions), _p(old_logp), _p(old_values), _p(adv), _p(sample_index),
                                _p(adv_stats3), branch_array(branch_sizes), len(branch_sizes), n, float(clip), float(beta),
                                float(vf_coef), _p(dlogits), _p(dvalue), _p(stats6), _p(scratch), _stream()), "trxl_ppo_loss")


def clip_adamw_step(params, grads, m, v, total, chunks, nchunks, ngroups, max_norm, lr, step, partial, norms,
                    betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    _check(load().trxl_clip_adamw_step(_p(params), _p(grads), _p(m), _p(v), int(total), _p(chunks), int(nchunks), int(ngroups),
                                       float(max_norm), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                       float(weight_decay), int(step), _p(partial), _p(norms), _stream()), "trxl_clip_adamw_step")


# ---- multi-GPU exchange ------------------------------------------------------------------------------------------
def comm_unique_id():
    """128 host bytes identifying a new NCCL clique (rank 0 creates it and ships it to the other ranks)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(load().trxl_comm_unique_id(buf), "trxl_comm_unique_id")
    return bytes(buf.raw)


def comm_create(id_bytes, rank, world_size):
    assert len(id_bytes) == COMM_ID_BYTES
    out = C.c_void_p(None)
    _check(load().trxl_comm_create(C.create_string_buffer(id_bytes, COMM_ID_BYTES), int(rank), int(world_size), C.byref(out)),
           "trxl_comm_create")
    return out.value


def comm_destroy(comm):
    if comm:
        load().trxl_comm_destroy(comm)


def comm_calls(comm):
    return int(load().trxl_comm_calls(comm))


def nccl_version():
    return int(load().trxl_comm_nccl_version())


def allreduce_grads(comm, tensor, count=None):
    """In-place fp32 sum of ``tensor[:count]`` over the ranks of ``comm`` on the current stream."""
    assert tensor.dtype == torch.float32
    _check(load().trxl_allreduce_grads(comm, _p(tensor), int(tensor.numel() if count is None else count), _stream()),
           "trxl_allreduce_grads")


def allreduce_f64(comm, tensor):
    assert tensor.dtype == torch.float64
    _check(load().trxl_allreduce_f64(comm, _p(tensor), int(tensor.numel()), _stream()), "trxl_allreduce_f64")
