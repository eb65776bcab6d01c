"""ctypes binding of libdynam3d_b200.so (the C ABI declared in include/dynam3d_b200.h).

There is NO fallback: if the shared library is missing or the device is not sm_100 the import of any
compute entry fails loudly (`D3DLibraryError`).  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynam3d_b200.so")

D3D_F16, D3D_BF16, D3D_OUT_F32 = 0, 1, 2
ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SILU, ACT_SWIGLU, ACT_LEAKY_RELU = 0, 1, 2, 3, 4, 5


class D3DLibraryError(RuntimeError):
    pass


class D3DError(RuntimeError):
    pass


class Mlp(ctypes.Structure):
    _fields_ = [("w0", ctypes.c_void_p), ("b0", ctypes.c_void_p), ("ln_g", ctypes.c_void_p), ("ln_b", ctypes.c_void_p),
                ("w3", ctypes.c_void_p), ("b3", ctypes.c_void_p),
                ("k_pad", ctypes.c_int), ("d_hidden", ctypes.c_int), ("d_out", ctypes.c_int), ("kind", ctypes.c_int)]


class EncoderLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("w_in", "b_in", "w_out", "b_out", "n1_g", "n1_b", "w1", "b1", "w2", "b2", "n2_g", "n2_b")]


class PoolLevel(ctypes.Structure):
    _fields_ = [("mlp", Mlp), ("agg", ctypes.c_void_p), ("layers", EncoderLayer * 2), ("norm_g", ctypes.c_void_p), ("norm_b", ctypes.c_void_p),
                ("norm_eps", ctypes.c_float), ("n_layers", ctypes.c_int), ("d_model", ctypes.c_int), ("n_head", ctypes.c_int)]


class FFPools(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("patch_pos", "patch_dir", "patch_scale", "patch_fts", "inst_pos", "inst_fts", "zone_pos", "zone_fts")]


class FFRuntime(ctypes.Structure):
    _fields_ = [("level_inst", ctypes.c_void_p), ("level_zone", ctypes.c_void_p), ("disc", ctypes.c_void_p),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
                ("stage_dev", ctypes.c_void_p), ("stage_host", ctypes.c_void_p), ("stage_bytes", ctypes.c_size_t),
                ("res_dev", ctypes.c_void_p), ("res_host", ctypes.c_void_p), ("d2", ctypes.c_void_p), ("idx", ctypes.c_void_p),
                ("disc_in", ctypes.c_void_p), ("disc_h32", ctypes.c_void_p), ("disc_h16", ctypes.c_void_p), ("disc_out", ctypes.c_void_p),
                ("out_merge", ctypes.c_void_p), ("out_zone", ctypes.c_void_p), ("event", ctypes.c_void_p), ("max_seq", ctypes.c_int)]


class ViTLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b", "w_fc", "b_fc", "w_pr", "b_pr")]


class ViTModel(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n_layers", "width", "n_heads", "patch", "tokens", "kpad", "resolution", "out_dim", "kind")] + \
               [("conv_w", ctypes.c_void_p), ("cls", ctypes.c_void_p), ("pos", ctypes.c_void_p), ("ln_pre_g", ctypes.c_void_p), ("ln_pre_b", ctypes.c_void_p),
                ("layers", ctypes.POINTER(ViTLayer)), ("ln_post_g", ctypes.c_void_p), ("ln_post_b", ctypes.c_void_p), ("proj", ctypes.c_void_p)]


class ViTScratch(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("cols", "conv", "X", "A16", "qkv", "att", "h", "out", "cu")]


class LMScratch(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("A16", "qkv", "att", "h", "rope_tab", "last16", "att_last", "x_last")]


class LMChunk(ctypes.Structure):
    _fields_ = [("seq_start", ctypes.c_void_p), ("seq_len", ctypes.c_void_p), ("rows", ctypes.c_void_p), ("positions", ctypes.c_void_p),
                ("q_tile_begin", ctypes.c_int), ("q_tile_end", ctypes.c_int)]


class LMLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("rms1", "w_qkv", "w_o", "rms2", "w_gu", "w_down")]


class LMModel(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int), ("hidden", ctypes.c_int), ("n_heads", ctypes.c_int), ("head_dim", ctypes.c_int), ("ffn", ctypes.c_int),
                ("vocab", ctypes.c_int), ("kind", ctypes.c_int), ("eps", ctypes.c_float), ("layers", ctypes.POINTER(LMLayer)),
                ("norm", ctypes.c_void_p), ("lm_head", ctypes.c_void_p), ("embed", ctypes.c_void_p)]


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64),
        ("W", ctypes.c_void_p), ("ldw", ctypes.c_int64),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_int64),
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("in_kind", ctypes.c_int), ("out_kind", ctypes.c_int),
        ("bias", ctypes.c_void_p), ("act", ctypes.c_int),
        ("residual", ctypes.c_void_p), ("ldres", ctypes.c_int64),
    ]


_lib = None

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_IP, _FP = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)

# argument types of every exported entry (mirrors include/dynam3d_b200.h); tests/test_abi.py checks the symbol list
SIGNATURES = {
    "d3d_version": [], "d3d_sm_count": [], "d3d_check_device": [_I], "d3d_launch_count": [],
    "d3d_gemm": [_P, _P], "d3d_gemm_simt": [_P, _P], "d3d_gemm_set_pair_mode": [_I],
    "d3d_depth_preprocess": [_P, _P, _I, _I, _I, _F, _F, _P],
    "d3d_depth_patch_grid": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _IP, _IP, _F, _F, _P],
    "d3d_unproject_habitat": [_P, _P, _I, _I, _I, _FP, _FP, _FP, _F, _P, _P, _P, _P],
    "d3d_patch_3d_info": [_P, _I, _I, _I, _FP, _FP, _FP, _F, _P, _P],
    "d3d_frustum_cull": [_P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _P, _F, _F, _F, _F, _F, _F, _F, _P, _P, _P],
    "d3d_frustum_cull_batched": [_P, _I, _I, _I, _P, _I, _I, _I, _P, _F, _F, _F, _F, _F, _F, _F, _P, _I, _P, _P],
    "d3d_frustum_cull_matrix": [_P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _P, _F, _F, _F, _P, _P, _P],
    "d3d_unproject_pinhole": [_P, _I, _I, _I, _P, _I, _I, _IP, _IP, _F, _F, _F, _P, _P, _P, _P, _P],
    "d3d_knn3d": [_P, _I, _P, _I, _I, _P, _P, _P],
    "d3d_seq_centroid": [_P, _P, _P, _I, _P, _P],
    "d3d_env_export": [_P, _P, _P, _I, _P, _F, _I, _P, _P, _P, _P],
    "d3d_env_export_batched": [_P, _P, _I, _I, _P, _P],
    "d3d_segm_relabel": [_P, _I, _I, _I, _I, _I, _I, _IP, _IP, _P, _P, _P],
    "d3d_layernorm": [_P, _L, _P, _P, _P, _F, _I, _I, _I, _P, _L, _P, _L, _I, _P],
    "d3d_rmsnorm": [_P, _L, _P, _P, _F, _I, _I, _P, _L, _P, _L, _I, _P],
    "d3d_rope": [_P, _L, _P, _P, _I, _I, _I, _I, _P],
    "d3d_gemm_skinny": [_P, _P], "d3d_gemm_skinny_set_config": [_I], "d3d_argmax_rows": [_P, _L, _I, _I, _P, _P],
    "d3d_decode_attention": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _P, _L, _P],
    "d3d_decode_attention_rope": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _L, _P],
    "d3d_lm_decode_set_pdl": [_I],
    "d3d_lm_decode_step": [_P, _P, _L, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "d3d_rope_table": [_P, _P, _I, _I, _P, _P], "d3d_rope_apply": [_P, _L, _P, _I, _I, _I, _I, _P],
    "d3d_embed_gather": [_P, _I, _P, _I, _I, _P, _L, _P],
    "d3d_preprocess_im2col": [_P, _I, _I, _I, _I, _I, _FP, _FP, _P, _I, _I, _P],
    "d3d_pil_resample_pass": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    "d3d_gather_rows16": [_P, _L, _P, _P, _L, _I, _I, _P],
    "d3d_vit_forward": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "d3d_phi3_prefill": [_P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _L, _P, _I, _P, _P],
    "d3d_scatter_rows16": [_P, _L, _P, _L, _P, _I, _I, _P],
    "d3d_attention_split_tc": [_P, _L, _L, _L, _P, _L, _P, _I, _I, _I, _I, _I, _F, _P],
    "d3d_wp_relu": [_P, _L, _P],
    "d3d_wp_neighbor_attention": [_P, _P, _I, _I, _I, _I, _F, _P, _P],
    "d3d_wp_heatmap_nms": [_P, _I, _I, _I, _I, _F, _F, _P, _P, _P],
    "d3d_attention_tc_set_halves": [_I, _I],
    "d3d_attention_tc_ex": [_P, _L, _L, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "d3d_phi3_prefill_chunk": [_P, _P, _I, _I, _I, _P, _P, _L, _L, _P, _P, _P, _P, _I, _P, _P],
    "d3d_gemm_set_sm_limit": [_I],
    "d3d_gemm_profile_begin": [], "d3d_gemm_profile_end": [_P, _P, _P],
    "d3d_vit_embed_ln": [_P, _P, _P, _P, _P, _F, _I, _I, _I, _P, _P],
    "d3d_scatter_rows": [_P, _L, _P, _P, _L, _P, _I, _I, _P],
    "d3d_add_inplace": [_P, _P, _L, _I, _P],
    "d3d_cast16": [_P, _L, _P, _L, _I, _I, _I, _P],
    "d3d_attention_simt": [_P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "d3d_attention_mma": [_P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "d3d_attention_mixed": [_P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _F, _P],
    "d3d_attention_tc": [_P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "d3d_pool_features": [_P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P],
    "d3d_pool_assemble": [_P, _P, _I, _P, _P, _P, _I, _I, _P, _P],
    "d3d_disc_input": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "d3d_patch_info_rows": [_P, _L, _P, _I, _P],
    "d3d_concat2_cast": [_P, _P, _I, _I, _P, _I, _P],
    "d3d_pos3_rows": [_P, _I, _P, _I, _P],
    "d3d_copy_blocks": [_P, _P, _P, _I, _P],
    "d3d_scatter_rows_ptr": [_P, _L, _P, _P, _I, _I, _P],
    "d3d_knn2_batched": [_P, _P, _P, _I, _P, _P, _P],
    "d3d_disc_input_batched": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "d3d_pool_workspace_bytes": [_I, _I, _I],
    "d3d_mlp_ln_gelu": [_P, _P, _L, _I, _P, _P, _P, _L, _P],
    "d3d_pool_tokens": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, ctypes.c_size_t, _P, _P],
    "d3d_ffh_create": [_I, _I, _F], "d3d_ffh_destroy": [_P], "d3d_ffh_reset": [_P, _I], "d3d_ffh_pop": [_P, _I],
    "d3d_ffh_counts": [_P, _I, _P], "d3d_ffh_cull": [_P, _I, _P, _L, _P, _P, _P, _P], "d3d_ffh_cull_list": [_P, _I, _P, _L, _P, _P, _P, _P], "d3d_ffh_set_tree": [_P],
    "d3d_ffh_begin_view": [_P] * 3 + [_I] + [_P] * 11,
    "d3d_ffh_begin_step": [_P] * 3 + [_I, _I] + [_P] * 9, "d3d_ffh_begin_view_refs": [_P, _I, _P],
    "d3d_ff_view_pre": [_P, _I, _P, _P, _P, _P, _P, _P, _P], "d3d_ff_view_post": [_P, _P, _P, _P, _P, _P], "d3d_ff_run_deferred": [_P, _P, _P],
    "d3d_vmm_create": [ctypes.c_size_t, ctypes.c_size_t, _I], "d3d_vmm_ensure": [_P, ctypes.c_size_t, _P], "d3d_vmm_base": [_P],
    "d3d_vmm_mapped": [_P], "d3d_vmm_reserved": [_P], "d3d_vmm_destroy": [_P],
    "d3d_event_create": [], "d3d_event_destroy": [_P], "d3d_ff_profile_begin": [], "d3d_ff_profile_end": [_P, _P, _P],
    "d3d_ffh_finish_view": [_P, _P, _P, _P], "d3d_ffh_fetch_view": [_P] * 17, "d3d_ffh_zone_key_array": [_P, _I, _P],
    "d3d_ffh_get_map": [_P, _I, _I, _P, _P, _P, _P], "d3d_ffh_get_p2i": [_P, _I, _P], "d3d_ffh_live_ids": [_P, _I, _I, _P, _P], "d3d_ffh_get_patch_pos": [_P, _I, _P],
    "d3d_ffh_get_zone_keys": [_P, _I, _P, _P, _P], "d3d_ffh_get_last": [_P, _I, _P, _P, _P, _P],
    "d3d_split16": [_P, _L, _P, _L, _I, _I, _I, _P],
    "d3d_attention_f32": [_P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _F, _P],
    "d3d_attention_split": [_P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _I, _F, _P],
    "d3d_ray_points_habitat": [_P, _P, _P, _I, _I] + [ctypes.c_double] * 5 + [_P, _P],
    "d3d_ray_topk": [_P, _P, _I, _I, _I, _F, _I, _P, _P],
    "d3d_gather_samples": [_P, _P, _I, _I, _I, _P, _P],
    "d3d_nerf_gather": [_P] * 8 + [_I, _I, _I, _I, _F, _F, _F, _F, _F, _P, _I, _P, _P],
    "d3d_add_half": [_P, _P, _P, _L, _P],
    "d3d_volume_render": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
}
OPTIONAL = set()


def _declare(lib_):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib_, name, None)
        if fn is None:
            if name in OPTIONAL:
                continue
            raise D3DLibraryError(f"{LIB_PATH} does not export {name}: stale build?")
        fn.argtypes = argtypes
        fn.restype = {"d3d_pool_workspace_bytes": ctypes.c_size_t, "d3d_ffh_create": ctypes.c_void_p, "d3d_ffh_destroy": None,
                      "d3d_event_create": ctypes.c_void_p, "d3d_event_destroy": None, "d3d_launch_count": ctypes.c_longlong,
                      "d3d_vmm_create": ctypes.c_void_p, "d3d_vmm_base": ctypes.c_uint64, "d3d_vmm_mapped": ctypes.c_size_t,
                      "d3d_vmm_reserved": ctypes.c_size_t}.get(name, ctypes.c_int)


def lib():
    """Load the library once; raise if it is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise D3DLibraryError(f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.d3d_last_error.restype = ctypes.c_char_p
        _declare(_lib)
    return _lib


N_CALLS = 0  # successful kernel-launching C-ABI calls (each launches >= 1 kernel); bench.py reports it as gpu_launches


def check(rc):
    global N_CALLS
    N_CALLS += 1
    if rc != 0:
        raise D3DError(f"libdynam3d_b200 error {rc}: {lib().d3d_last_error().decode()}")


def require_device(dev=0):
    import torch
    if not torch.cuda.is_available():
        raise D3DLibraryError("no CUDA device: dynam3d_b200 has no CPU fallback")
    check(lib().d3d_check_device(int(dev)))


_STREAM = None  # cached stream handle inside `stream_scope()` (torch.cuda.current_stream() costs ~15 us per query)


def stream_ptr():
    if _STREAM is not None:
        return _STREAM
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class stream_scope:
    """Caches the current stream for the duration of a host-side hot loop; nests; do not switch streams inside."""

    def __enter__(self):
        global _STREAM
        import torch
        self.prev = _STREAM
        _STREAM = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # re-query: scopes may nest inside torch.cuda.stream(...)
        return self

    def __exit__(self, *a):
        global _STREAM
        _STREAM = self.prev
        return False


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def kind_of(dtype):
    import torch
    if dtype == torch.float16:
        return D3D_F16
    if dtype == torch.bfloat16:
        return D3D_BF16
    if dtype == torch.float32:
        return D3D_OUT_F32
    raise TypeError(dtype)
