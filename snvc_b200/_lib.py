"""ctypes binding of libsnvc_b200.so (the C ABI declared in include/snvc_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised
(the reference raises through AT_ASSERTM / AT_ERROR, BuildCostVolume_cuda.cu:212-220)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnvc_b200.so")

F32, BF16, F64 = 0, 1, 2
NCDHW, NDHWC = 0, 1

_i64, _i32, _p, _f = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_float


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, _i32) for n in (
        "N", "Cin", "Cout", "Di", "Hi", "Wi", "Do", "Ho", "Wo", "kernel", "stride", "pad", "dilation",
        "transposed", "relu", "residual_mode", "sigmoid", "out_dtype", "out_cstride", "out_coffset",
        "res_cstride", "res_coffset", "in_cstride", "in_coffset", "addend_edge_lo", "addend_edge_hi")]


class Conv2dDesc(ctypes.Structure):
    _fields_ = [(n, _i32) for n in (
        "N", "Cin", "Cout", "Hi", "Wi", "Ho", "Wo", "kernel", "stride", "pad", "dilation", "transposed", "relu",
        "residual_mode", "sigmoid", "out_dtype", "out_cstride", "out_coffset", "res_cstride", "res_coffset",
        "in_cstride", "in_coffset")] + [("reserved", _i32 * 2)]


_SIGS = {
    "snvc_version": ([], _i32),
    "snvc_last_error": ([], ctypes.c_char_p),
    "snvc_launch_count": ([], _i64),
    "snvc_set_option": ([ctypes.c_char_p, ctypes.c_char_p], _i32),
    "snvc_get_option": ([ctypes.c_char_p], ctypes.c_char_p),
    "snvc_cost_volume_fwd": ([_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _p], _i32),
    "snvc_cost_volume_split_fwd": ([_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _p], _i32),
    "snvc_cost_volume_bwd": ([_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p], _i32),
    "snvc_cost_volume_xlow": ([_p, _p, _i64, _i64, _i64, _i32, _p], _i32),
    "snvc_roi_voxel_sample_workspace_bytes": ([_i64, _i64, _i64, _i64], _i64),
    "snvc_roi_voxel_sample_fwd": ([_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _f, _f, _i32, _i32, _p], _i32),
    "snvc_roi_voxel_sample_indices": ([_p, _p, _p, _i64, _i64, _i64, _i64, _f, _f, _p], _i32),
    "snvc_frustum_lift_fwd": ([_p, _p, _p, _p, _p, ctypes.POINTER(_f), _p, _p] + [_i64] * 8 + [_i32] * 5 + [_p], _i32),
    "snvc_frustum_lift_slab_fwd": ([_p, _p, _p, _p, _p, ctypes.POINTER(_f), _p, _p] + [_i64] * 8 + [_i32] * 5 + [_i64, _i64, _p],
                                   _i32),
    "snvc_frustum_lift_indices": ([_p, _p, _p, _p, ctypes.POINTER(_f), _p, _p] + [_i64] * 7 + [_i32, _p], _i32),
    "snvc_roi_grid_project": ([_p] * 11 + [_i64] * 4 + [_p], _i32),
    "snvc_disparity_regression": ([_p, _p, _p, _i64, _i64, _i64, _p], _i32),
    "snvc_depth_regression_fwd": ([_p, _p, _p] + [_i64] * 7 + [_i32, _p], _i32),
    "snvc_boxes_iou_bev": ([_p, _p, _p, _i64, _i64, _p], _i32),
    "snvc_nms_bev_workspace_bytes": ([_i64], _i64),
    "snvc_nms_bev": ([_p, _p, _p, _p, _i64, _f, _p], _i32),
    "snvc_nms_bev_batched": ([_p, _p, _p, _p, _i64, _i64, _f, _p], _i32),
    "snvc_conv3d_packed_weight_bytes": ([_i32, _i32, _i32], _i64),
    "snvc_conv3d_pack_weights": ([_p, _p, _i32, _i32, _i32, _i32, _p], _i32),
    "snvc_conv3d_fwd": ([_p, _p, _p, _p, _p, _p, ctypes.POINTER(ConvDesc), _p], _i32),
    "snvc_conv3d_fwd_addend": ([_p, _p, _p, _p, _p, _p, ctypes.POINTER(ConvDesc), _p], _i32),
    "snvc_group_norm_workspace_bytes": ([_i64, _i64, _i32], _i64),
    "snvc_group_norm_fwd": ([_p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _f] + [_i32] * 8 + [_p], _i32),
    "snvc_conv2d_packed_weight_bytes": ([_i32, _i32, _i32], _i64),
    "snvc_conv2d_pack_weights": ([_p, _p, _i32, _i32, _i32, _i32, _p], _i32),
    "snvc_conv2d_fwd": ([_p, _p, _p, _p, _p, _p, ctypes.POINTER(Conv2dDesc), _p], _i32),
    "snvc_ncdhw_f32_to_ndhwc_bf16": ([_p, _p, _i64, _i64, _i64, _p], _i32),
    "snvc_ndhwc_bf16_to_ncdhw_f32": ([_p, _p, _i64, _i64, _i64, _p], _i32),
    "snvc_scale_by_occupancy": ([_p, _p, _p, _i64, _i32, _i32, _i32, _p], _i32),
    "snvc_avgpool_to_bev": ([_p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p], _i32),
    "snvc_avgpool_to_bev_nhwc": ([_p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _p], _i32),
    "snvc_halo_unique_id": ([_p], _i32),
    "snvc_halo_comm_create": ([_p, _i32, _i32, ctypes.POINTER(ctypes.c_void_p)], _i32),
    "snvc_halo_comm_destroy": ([_p], _i32),
    "snvc_halo_exchange": ([_p, _p, _i64, _i64, _i32, _i32, _i32, _p], _i32),
    "snvc_peer_alloc": ([_i64, ctypes.POINTER(ctypes.c_void_p)], _i32),
    "snvc_peer_free": ([_p], _i32),
    "snvc_peer_export": ([_p, _p], _i32),
    "snvc_peer_open": ([_p, ctypes.POINTER(ctypes.c_void_p)], _i32),
    "snvc_peer_close": ([_p], _i32),
    "snvc_peer_ctl_bytes": ([], _i64),
    "snvc_halo_push": ([_p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _i32, _p], _i32),
    "snvc_masked_rows_to_host": ([_p, _p, _p, _p, _i64, _i32, _i32, _p, _p], _i32),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -m snvc_b200.build`); snvc_b200 has no CPU fallback")
        l = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            f = getattr(l, name)
            f.argtypes, f.restype = args, res
        _lib = l
    return _lib


def set_option(name, value=None):
    """Debug / A-B switch of the library (see include/snvc_b200.h); value None = unset, name None = unset all."""
    st = lib().snvc_set_option(name.encode() if name is not None else None, str(value).encode() if value is not None else None)
    check(st, "snvc_set_option")


def get_option(name):
    v = lib().snvc_get_option(name.encode())
    return v.decode() if v else None


def check(status, what):
    if status != 0:
        msg = lib().snvc_last_error()
        raise RuntimeError(f"{what} failed ({status}): {msg.decode() if msg else ''}")


def stream_ptr():
    """cudaStream_t of torch's CURRENT stream on the current device (reference: .cu:230)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU")   # BuildCostVolume.cpp:26
