"""ctypes binding of libdac_b200.so (include/dac_b200.h).  No compute happens in Python: tensors are only
device memory here; every op is a C-ABI call that enqueues hand-written sm_100a kernels on the current
CUDA stream.  There is no CPU or PyTorch-op fallback - if the library or a GPU is missing, calls raise."""
import ctypes as C
import os

import torch

from . import build as _build

_lib = None


class DacError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("c0", C.c_int32), ("ld0", C.c_int32),
        ("src1", C.c_void_p), ("c1", C.c_int32), ("ld1", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("OH", C.c_int32), ("OW", C.c_int32),
        ("stride", C.c_int32), ("ngroups", C.c_int32), ("ntaps", C.c_int32),
        ("ndy", C.c_int32), ("ncols", C.c_int32),
        ("col_dx", (C.c_int8 * 16) * 4), ("col_dy0", (C.c_int8 * 16) * 4), ("col_tap", (C.c_int8 * 16) * 4),
        ("out_scale", C.c_int32), ("out_oy", C.c_int8 * 4), ("out_ox", C.c_int8 * 4),
        ("weight", C.c_void_p), ("cout", C.c_int32), ("cout_pad", C.c_int32), ("per_image_w", C.c_int32),
        ("block_n", C.c_int32),
        ("rsrc0", C.c_void_p), ("rc0", C.c_int32), ("rld0", C.c_int32),
        ("rsrc1", C.c_void_p), ("rc1", C.c_int32), ("rld1", C.c_int32),
        ("rweight", C.c_void_p),
        ("tile_h", C.c_int32), ("tile_w", C.c_int32),
        ("epi", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p), ("bias_img", C.c_void_p),
        ("film", C.c_void_p), ("film_ld", C.c_int32), ("film_off", C.c_int32),
        ("ln_g", C.c_void_p), ("ln_eps", C.c_float),
        ("res", C.c_void_p), ("res_ld", C.c_int32),
        ("res2", C.c_void_p), ("res2_ld", C.c_int32),
        ("out", C.c_void_p), ("out_ld", C.c_int32), ("out_coff", C.c_int32),
        ("res_f32", C.c_void_p), ("res_f32_ld", C.c_int32),
        ("out_f32", C.c_void_p), ("out_f32_ld", C.c_int32),
        ("out_planar", C.c_void_p),
        ("stats_out", C.c_void_p), ("stats_eps", C.c_float),
        ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p),
        ("out_nchw", C.c_void_p), ("out_nchw_c", C.c_int32), ("out_nchw_h", C.c_int32), ("out_nchw_w", C.c_int32),
        ("kv_shift", C.c_void_p), ("ctx_acc", C.c_void_p),
        ("halo", C.c_int32),
        ("ctx_slots", C.c_int32),
        ("pair", C.c_int32),
    ]


class EmbedWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "time_w1", "time_b1", "time_w2", "time_b2", "text_w1", "text_b1", "text_w2", "text_b2",
        "prompt", "prompt_w", "prompt_b", "film_w", "film_b")] + [
        ("nf", C.c_int32), ("time_dim", C.c_int32), ("ctx_dim", C.c_int32), ("F", C.c_int32)]


EPI_PLAIN, EPI_GEGLU, EPI_LN, EPI_QKV, EPI_KVCTX = 0, 1, 2, 3, 4
ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2

# every symbol include/dac_b200.h declares: (restype, argtypes)
_i32, _i64, _f, _p = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = {
    "dac_version": (C.c_int, []),
    "dac_last_error": (C.c_char_p, []),
    "dac_launch_count": (_i64, []),
    "dac_reset_launch_count": (None, []),
    "dac_abi_sizes": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dac_sde_step": (C.c_int, [C.c_int, _p, _p, _p, _p, _p, _i64, C.POINTER(C.c_float), _p]),
    "dac_loop_tick": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "dac_sde_step_dev": (C.c_int, [C.c_int, _p, _p, _p, _p, C.c_int64, _p, _p, _p]),
    "dac_noise_state": (C.c_int, [_p, _p, _p, _i64, _f, _p]),
    "dac_unet_stem_input": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32, _p]),
    "dac_conv_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_p)]),
    "dac_conv_launch": (C.c_int, [_p, _p]),
    "dac_conv_destroy": (None, [_p]),
    "dac_conv_info": (C.c_int, [_p] + [C.POINTER(_i32)] * 4),
    "dac_layernorm_rows": (C.c_int, [_p, _i32, _p, _i32, _i64, _i32, _p, _p, _f, _p]),
    "dac_layernorm_rows_res": (C.c_int, [_p, _i32, _p, _i32, _p, _i32, _i64, _i32, _p, _p, _f, _p]),
    "dac_layernorm_rows_f32": (C.c_int, [_p, _i32, _p, _i32, _i64, _i32, _p, _p, _f, _p]),
    "dac_groupnorm_nhwc": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p, _p, _f, _p, _p]),
    "dac_prenorm_groupnorm_nhwc": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _p, _f, _p, _p, _f, _p, _p]),
    "dac_prompt_embed": (C.c_int, [C.POINTER(EmbedWeights), _p, _i32, _p, _p]),
    "dac_time_film": (C.c_int, [C.POINTER(EmbedWeights), _p, _p, _i32, _p, _p, _p]),
    "dac_two_linear": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _p, _i32, _p, _p]),
    "dac_linattn_context": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "dac_linattn_fold": (C.c_int, [_p, _i32, _i32, _i32, _p, _i32, _i32, _p, _p]),
    "dac_linattn_fold_g": (C.c_int, [_p, _i32, _i32, _i32, _p, _i32, _i32, _p, _p]),
    "dac_attention": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p]),
    "dac_vit_patchify": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p]),
    "dac_vit_embed": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _f, _p]),
    "dac_vit_pool": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _f, _p, _i32, _p, _p]),
    "dac_set_pdl": (None, [_i32]),
    "dac_attention_causal": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p]),
    "dac_text_embed": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p]),
    "dac_text_pool": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _f, _p, _i32, _p, _p]),
    "dac_degradation_argmax": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "dac_linattn_ctx_slots": (C.c_int32, [_i32, _i32, _i32]),
    "dac_linattn_kv_create": (C.c_int, [_p, _p, _p, _p, _i32, _p, _p, _i32, _i32, _i32, _i32, _f, C.POINTER(_p)]),
    "dac_linattn_kv_launch": (C.c_int, [_p, _p]),
    "dac_linattn_kv_destroy": (None, [_p]),
    "dac_linattn_qout_create": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _f, _p, _p, _i32, _i32, _i32, _i32, _f, _p, C.POINTER(_p)]),
    "dac_linattn_qout_launch": (C.c_int, [_p, _p]),
    "dac_linattn_qout_destroy": (None, [_p]),
    "dac_clip_resample_h": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _p, _i32, _i32, _i32, _p]),
    "dac_clip_resample_v_norm": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _i32, _i32, _i32,
                                           C.POINTER(C.c_float), C.POINTER(C.c_float), _p, _p]),
    "dac_tensor2img": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _f, _f, _p]),
}


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Loads (building first if needed) the C-ABI library and types every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("DAC_DEBUG") == "1":
        path = _build.build(debug=True)
    elif os.environ.get("DAC_LIB"):            # a hand-built variant of the library (kernel debugging)
        path, build_if_missing = os.environ["DAC_LIB"], False
    if build_if_missing:
        _build.build()       # no-op when the library matches the content hash of csrc/ (never runs a stale binary)
    if not os.path.exists(path):
        raise DacError(f"{path} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    a, b = C.c_int32(), C.c_int32()
    lib.dac_abi_sizes(C.byref(a), C.byref(b))
    if (a.value, b.value) != (C.sizeof(ConvDesc), C.sizeof(EmbedWeights)):
        raise DacError(f"{path} is stale: struct sizes {a.value},{b.value} != binding "
                       f"{C.sizeof(ConvDesc)},{C.sizeof(EmbedWeights)}; rebuild with `python da-clip_b200/build.py --force`")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DacError(f"dac error {rc}: {load().dac_last_error().decode()}")


def require_cuda(*tensors):
    """Every tensor handed to a kernel must live on the CURRENT CUDA device: launches, TMA descriptor encodes and the
    stream from stream_ptr() all target it.  Entry points switch to their tensors' device first (on_device)."""
    if not torch.cuda.is_available():
        raise DacError("daclip_b200 needs a CUDA device: there is no CPU fallback for this path")
    cur = torch.cuda.current_device()
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DacError("daclip_b200 kernels take CUDA tensors only (no CPU fallback)")
        if t.device.index != cur:
            raise DacError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: "
                           f"wrap the call in `with torch.cuda.device({t.device.index})`")


def on_device(dev):
    """Context manager making `dev` (a torch.device / tensor / index) the current CUDA device, as PyTorch ops do
    implicitly for their arguments; a no-op for the device that is already current."""
    if torch.is_tensor(dev):
        dev = dev.device
    return torch.cuda.device(dev)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def set_pdl(on):
    load().dac_set_pdl(1 if on else 0)


def launch_count():
    return int(load().dac_launch_count())


def reset_launch_count():
    load().dac_reset_launch_count()
