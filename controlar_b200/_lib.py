"""ctypes binding of include/controlar_b200.h.  There is NO fallback: if the CUDA library is missing or a call
fails, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAR_LIB") or os.path.join(HERE, "lib", "libcontrolar_b200.so")   # CAR_LIB: dev A/B runs of another build

CAR_BF16, CAR_F32 = 0, 1


class CarModelDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("dim", C.c_int32), ("n_layer", C.c_int32), ("n_head", C.c_int32),
                ("ffn_dim", C.c_int32), ("vocab_size", C.c_int32), ("cls_token_num", C.c_int32),
                ("block_size", C.c_int32), ("caption_dim", C.c_int32), ("model_type", C.c_int32),
                ("norm_eps", C.c_float), ("rope_base", C.c_float)]


class CarWeights(C.Structure):
    _fields_ = [("tok_embeddings", C.c_void_p), ("norm", C.c_void_p), ("output", C.c_void_p),
                ("attention_norm", C.POINTER(C.c_void_p)), ("wqkv", C.POINTER(C.c_void_p)),
                ("wo", C.POINTER(C.c_void_p)), ("ffn_norm", C.POINTER(C.c_void_p)),
                ("w1", C.POINTER(C.c_void_p)), ("w3", C.POINTER(C.c_void_p)), ("w2", C.POINTER(C.c_void_p)),
                ("cap_fc1", C.c_void_p), ("cap_fc2", C.c_void_p), ("label_table", C.c_void_p),
                ("cond_fc1", C.c_void_p), ("cond_fc2", C.c_void_p),
                ("ctl_fc1", C.c_void_p * 3), ("ctl_fc2", C.c_void_p * 3)]


class CarTrainWeights(C.Structure):
    _fields_ = [("w", CarWeights), ("adapter_fc1", C.c_void_p), ("adapter_fc2", C.c_void_p), ("cap_uncond", C.c_void_p),
                ("adapter_dim", C.c_int32), ("num_classes", C.c_int32), ("cond_uncond", C.c_void_p)]


class CarSampling(C.Structure):
    _fields_ = [("temperature", C.c_float), ("top_k", C.c_int32), ("top_p", C.c_float),
                ("sample_logits", C.c_int32), ("cfg_scale", C.c_float), ("cfg_interval", C.c_int32),
                ("seed", C.c_uint64)]


# name -> (restype, argtypes); every symbol declared in include/controlar_b200.h
PROTOTYPES = {
    "car_last_error": (C.c_char_p, []),
    "car_version": (C.c_int, []),
    "car_model_create": (C.c_int, [C.POINTER(CarModelDesc), C.POINTER(CarWeights), C.c_void_p, C.POINTER(C.c_void_p)]),
    "car_model_repack": (C.c_int, [C.c_void_p, C.POINTER(CarWeights), C.c_void_p]),
    "car_model_destroy": (C.c_int, [C.c_void_p]),
    "car_state_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_void_p)]),
    "car_state_set_emb_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_state_destroy": (C.c_int, [C.c_void_p]),
    "car_state_set_step_timer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "car_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32, C.c_void_p]),
    "car_decode_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "car_sample": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(CarSampling), C.c_int32, C.c_int32,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_generate": (C.c_int, [C.c_void_p, C.POINTER(CarSampling), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_generate_forced": (C.c_int, [C.c_void_p, C.POINTER(CarSampling), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "car_decode_step_bytes": (C.c_int64, [C.c_void_p, C.c_int32]),
    "car_launch_count": (C.c_int64, [C.c_int32]),
    "car_op_linear": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                C.c_int32, C.c_int32, C.c_void_p]),
    "car_op_dense_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "car_train_create": (C.c_int, [C.POINTER(CarModelDesc), C.POINTER(CarTrainWeights), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_void_p)]),
    "car_train_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_train_backward": (C.c_int, [C.c_void_p, C.POINTER(CarTrainWeights), C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_train_destroy": (C.c_int, [C.c_void_p]),
    "car_canny_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "car_canny_u8": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_void_p, C.c_void_p]),
    "car_left_pad_captions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_hed_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "car_hed_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "car_hed_destroy": (C.c_int, [C.c_void_p]),
    "car_t5_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "car_t5_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "car_t5_destroy": (C.c_int, [C.c_void_p]),
    "car_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "car_op_rmsnorm": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                                 C.c_void_p]),
}

_lib = None


def lib():
    """Load the library (once).  Raises if it has not been built — there is no CPU or eager fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m controlar_b200.build` "
                "(controlar_b200 has no CPU / eager-PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            if os.environ.get("CAR_LIB") and not hasattr(l, name):
                continue                                  # an older dev build may lack the newest entry points
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().car_last_error()
        raise RuntimeError(f"controlar_b200 {what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def dtype_code(dt) -> int:
    import torch
    if dt == torch.bfloat16:
        return CAR_BF16
    if dt == torch.float32:
        return CAR_F32
    raise RuntimeError(f"controlar_b200 supports bf16 and fp32 checkpoints, not {dt}")


def cur_stream(device=None) -> int:
    """Raw handle of torch's current stream on `device` (default: the current device)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def on_device(t):
    """Context manager that makes the device of tensor `t` current for the duration of a library call: the library launches on
    the current device and takes the stream from it (the reference wraps its calls in `with torch.device(device)`,
    generate.py:179-182)."""
    import torch
    return torch.cuda.device(t.device)


def _no_handle():
    return None


class NativeHandle:
    """Base of the objects that own a library handle.  Handles are never copied: `copy.deepcopy` / pickling of the owning module
    (e.g. `ema = deepcopy(model)` of the train scripts, train_c2i_canny.py:117, after the model has been used) yields None in the
    copy, which rebuilds its own handle lazily on first use — instead of ctypes' "objects containing pointers cannot be pickled"."""

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_handle, ())


def _ptr(t):
    """Device pointer of a contiguous CUDA tensor (None passes through)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("controlar_b200: tensor is not on a CUDA device (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("controlar_b200: tensor must be contiguous")
    return t.data_ptr()


def _ptr_array(ts):
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = _ptr(t)
    return arr


from . import vision as _vision  # noqa: E402,F401  (registers the car_dino_* / car_vq_* prototypes in PROTOTYPES)
