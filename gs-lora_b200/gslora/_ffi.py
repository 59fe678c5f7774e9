"""ctypes binding of libgslora.so (include/gslora.h).  The library is built in-tree by
`__graft_entry__.build()` / `make -C gs-lora_b200/csrc`; there is NO fallback: if the shared object is
missing or a call fails, an exception is raised."""
import ctypes
import os

import torch

_LIB = None
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libgslora.so")

c_void_p, c_int, c_int64, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t


class GslError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GslError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(gslora-b200 has no CPU / PyTorch fallback)")
        _LIB = ctypes.CDLL(LIB_PATH)
        _declare(_LIB)
    return _LIB


class GslConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("image_size", "patch_size", "channels", "dim", "depth", "heads", "mlp_dim", "num_class",
                                              "lora_rank", "max_batch", "num_slots", "patch_order")] + \
               [(n, ctypes.c_float) for n in ("attn_scale", "ln_eps", "cos_s", "cos_m", "lora_scaling", "grad_scale", "dropout", "emb_dropout")] + \
               [("head_type", ctypes.c_int32), ("precision", ctypes.c_int32), ("lora_pos", ctypes.c_int32)]


PRECISION_FAST, PRECISION_SPLIT, PRECISION_SPLIT8 = 0, 1, 2
PRECISION_BY_NAME = {"fast": PRECISION_FAST, "split": PRECISION_SPLIT, "split8": PRECISION_SPLIT8}


def default_precision() -> int:
    """GSLORA_PRECISION = split8 (default: fp16 hi + e4m3 lo weights, the lo term on the FP8 tensor path; LoRA gradients within the 1e-3 parity
    bar) | split (fp16 hi + fp16 lo: two fp16 MMAs per k-step, same accuracy, 5 % slower) | fast (one fp16 rounding per weight: misses the bar)."""
    v = os.environ.get("GSLORA_PRECISION", "split8").lower()
    if v not in PRECISION_BY_NAME:
        raise GslError(f"GSLORA_PRECISION={v!r}: expected one of {sorted(PRECISION_BY_NAME)}")
    return PRECISION_BY_NAME[v]


SLOT_EMB, SLOT_LOGITS, SLOT_CE, SLOT_CORRECT, SLOT_XFINAL = range(5)
NUM_GLOBAL_PARAMS, NUM_BLOCK_PARAMS = 8, 12


def _declare(L):
    L.gsl_last_error.restype = ctypes.c_char_p
    L.gsl_version.restype = c_int
    L.gsl_launch_count.restype = ctypes.c_longlong
    L.gsl_set_gemm_cta_group.argtypes = [c_int]
    L.gsl_gemm_f16.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p,
                               c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_float, ctypes.c_uint32, c_void_p]
    L.gsl_gemm_f16.restype = c_int
    L.gsl_gemm_f16_split.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p,
                                     c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_float, ctypes.c_uint32, c_void_p]
    L.gsl_gemm_f16_split.restype = c_int
    L.gsl_gemm_f16_split8.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p,
                                      c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_float, ctypes.c_uint32, c_void_p]
    L.gsl_gemm_f16_split8.restype = c_int
    P = c_void_p
    sigs = {
        "gsl_patchify_f16": [P, P, c_int64, c_int, c_int, c_int, c_int, c_int, P],
        "gsl_patchify_u8_f16": [P, c_int, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, P],
        "gsl_engine_forward_u8": [P, c_int, P, c_int, P, P, P, c_int, c_int, ctypes.c_uint64, P],
        "gsl_class_sums": [P, P, c_int, c_int, c_int, P, P, P],
        "gsl_class_means": [P, P, c_int, c_int, P, P],
        "gsl_layernorm_fwd": [P, c_int64, P, P, c_float, P, c_int64, P, P, c_int64, c_int, P],
        "gsl_layernorm_bwd": [P, c_int64, P, c_int64, P, P, P, P, c_int64, P, c_int64, P, c_int64, c_int64, c_int, P],
        "gsl_lora_down": [P, c_int64, P, c_int64, P, c_int64, c_int64, c_int, c_int, P],
        "gsl_lora_down_split": [P, c_int64, P, c_int64, P, c_int64, c_int64, c_int, c_int, P],
        "gsl_lora_side_split": [P, c_int64, P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int, c_float, c_int, c_int64, c_int, c_int, P, c_size_t, P],
        "gsl_cast_f32_to_f16_split": [P, c_int64, P, P, c_int64, c_int64, c_int64, c_float, c_int, P],
        "gsl_cast_f32_to_f16_split8": [P, c_int64, P, P, c_int64, c_int64, c_int64, c_int, c_int, P],
        "gsl_skinny_tn": [P, c_int64, P, c_int64, P, c_int64, c_int, c_float, c_int, c_int64, c_int, c_int, P, c_size_t, P],
        "gsl_lora_side": [P, c_int64, P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int, c_float, c_int, c_int64, c_int, c_int, P, c_size_t, P],
        "gsl_attention_fwd": [P, c_int64, P, c_int64, P, c_int, c_int, c_int, c_float, P],
        "gsl_attention_bwd": [P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, c_int, c_int, c_float, P],
        "gsl_attention_bwd_rowdot": [P, c_int64, P, c_int64, P, P, P, c_int64, c_int, c_int, c_int, c_float, P],
        "gsl_cast_f32_to_f16": [P, c_int64, P, c_int64, c_int64, c_int64, c_float, c_int, P],
        "gsl_grouplasso_adamw_step": [P, P, P, P, P, c_int, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, c_float, c_int, P, P],
        "gsl_tensor_norms": [P, P, c_int, c_int, P, P],
        "gsl_engine_create": [ctypes.POINTER(GslConfig), P, c_size_t, ctypes.POINTER(P)],
        "gsl_engine_bind_params": [P, ctypes.POINTER(P), c_int, P, P],
        "gsl_engine_refresh_frozen": [P, P],
        "gsl_engine_refresh_lora": [P, P],
        "gsl_engine_forward": [P, c_int, P, P, c_int, c_int, ctypes.c_uint64, P],
        "gsl_engine_backward": [P, c_int, P, P, c_int, P],
        "gsl_engine_forward_dev": [P, c_int, P, c_int, P, c_int, c_int, c_int, P, P],
        "gsl_grouplasso_adamw_step_dev": [P, P, P, P, P, c_int, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, P, P, P],
        "gsl_loss_sums": [P, P, P, c_int, c_int, P, P],
        "gsl_prototype_kl_fwd": [P, P, P, c_int, c_int, P, P],
        "gsl_prototype_kl_grad": [P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_float, P, P],
        "gsl_unlearn_ce_grad": [P, P, P, c_int, c_int, c_int, c_float, c_float, P, P],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = c_int
    L.gsl_skinny_tn_workspace.argtypes = [c_int64, c_int, c_int]
    L.gsl_skinny_tn_workspace.restype = c_size_t
    L.gsl_lora_side_workspace.argtypes = [c_int64, c_int, c_int]
    L.gsl_lora_side_workspace.restype = c_size_t
    L.gsl_engine_workspace_bytes.argtypes = [ctypes.POINTER(GslConfig)]
    L.gsl_engine_workspace_bytes.restype = c_size_t
    L.gsl_engine_destroy.argtypes = [P]
    L.gsl_engine_destroy.restype = None
    L.gsl_engine_slot_ptr.argtypes = [P, c_int, c_int]
    L.gsl_engine_slot_ptr.restype = P
    L.gsl_engine_lora_offset.argtypes = [P, c_int, c_int]
    L.gsl_engine_lora_offset.restype = c_int64
    L.gsl_engine_lora_numel.argtypes = [P]
    L.gsl_engine_lora_numel.restype = c_int64
    L.gsl_count_launches.argtypes = [ctypes.c_longlong]
    L.gsl_count_launches.restype = None


EXPORTS = ["gsl_last_error", "gsl_version", "gsl_launch_count", "gsl_set_gemm_cta_group", "gsl_gemm_f16", "gsl_patchify_f16", "gsl_layernorm_fwd",
           "gsl_layernorm_bwd", "gsl_lora_down", "gsl_skinny_tn_workspace", "gsl_skinny_tn", "gsl_lora_side_workspace", "gsl_lora_side", "gsl_attention_fwd", "gsl_attention_bwd", "gsl_attention_bwd_rowdot",
           "gsl_cast_f32_to_f16", "gsl_grouplasso_adamw_step", "gsl_tensor_norms", "gsl_engine_workspace_bytes", "gsl_engine_create",
           "gsl_engine_destroy", "gsl_engine_bind_params", "gsl_engine_refresh_frozen", "gsl_engine_refresh_lora", "gsl_engine_forward",
           "gsl_engine_backward", "gsl_engine_slot_ptr", "gsl_engine_lora_offset", "gsl_engine_lora_numel", "gsl_loss_sums",
           "gsl_unlearn_ce_grad", "gsl_prototype_kl_fwd", "gsl_prototype_kl_grad", "gsl_patchify_u8_f16", "gsl_engine_forward_u8",
           "gsl_class_sums", "gsl_class_means", "gsl_gemm_f16_split", "gsl_lora_down_split", "gsl_lora_side_split", "gsl_cast_f32_to_f16_split", "gsl_cast_f32_to_f16_split8", "gsl_gemm_f16_split8",
           "gsl_engine_forward_dev", "gsl_grouplasso_adamw_step_dev", "gsl_count_launches"]


def host_floats(values):
    """ctypes float array for the HOST-pointer arguments (mean / std of gsl_patchify_u8_f16), or None."""
    if values is None:
        return None
    vals = [float(v) for v in values]
    return (ctypes.c_float * len(vals))(*vals)


def check(rc, what=""):
    if rc != 0:
        msg = lib().gsl_last_error().decode(errors="replace")
        raise GslError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "gslora-b200 operates on CUDA tensors only (no CPU fallback)"
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


EPI_F16, EPI_F32, EPI_GELU, EPI_GELU_BWD, EPI_RES_F32, EPI_PERIODIC_F32, EPI_F16_ROWDOT = range(7)


def gemm_f16(A, B, *, epi=EPI_F16, bias=None, out0, out1=None, aux=None, aux_period=0, K=None, N=None, M=None,
             cta_group=0, block_n=0, drop_p=0.0, drop_seed=0, B_lo=None, B_lo8=None, lo8_shift=0):
    """out = epi(A[:, :K] @ (B [+ B_lo])[:, :K].T); A, B fp16 row-major 2-D (possibly column-sliced views); B_lo: split-weight residual."""
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = B.shape[0] if N is None else N
    if B_lo8 is not None:       # split8: B = fp16(W * 2^shift), B_lo8 = e4m3 residual bytes (cast_split8)
        assert B_lo8.shape == B.shape and B_lo8.stride(0) == B.stride(0) and B_lo8.dtype == torch.uint8
        rc = lib().gsl_gemm_f16_split8(ptr(A), A.stride(0), ptr(B), ptr(B_lo8), int(lo8_shift), B.stride(0), M, N, K, epi, ptr(bias),
                                       ptr(out0), out0.stride(0), ptr(out1), out1.stride(0) if out1 is not None else 0,
                                       ptr(aux), aux.stride(0) if aux is not None else 0, aux_period, block_n, float(drop_p), int(drop_seed),
                                       cur_stream())
        check(rc, "gsl_gemm_f16_split8")
        return
    if B_lo is not None:
        assert B_lo.shape == B.shape and B_lo.stride(0) == B.stride(0)
        rc = lib().gsl_gemm_f16_split(ptr(A), A.stride(0), ptr(B), ptr(B_lo), B.stride(0), M, N, K, epi, ptr(bias),
                                      ptr(out0), out0.stride(0), ptr(out1), out1.stride(0) if out1 is not None else 0,
                                      ptr(aux), aux.stride(0) if aux is not None else 0, aux_period, cta_group, block_n, float(drop_p),
                                      int(drop_seed), cur_stream())
        check(rc, "gsl_gemm_f16_split")
        return
    rc = lib().gsl_gemm_f16(ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, epi, ptr(bias),
                            ptr(out0), out0.stride(0), ptr(out1), out1.stride(0) if out1 is not None else 0,
                            ptr(aux), aux.stride(0) if aux is not None else 0, aux_period, cta_group, block_n, float(drop_p), int(drop_seed),
                            cur_stream())
    check(rc, "gsl_gemm_f16")
