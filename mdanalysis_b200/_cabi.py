"""ctypes binding of libmdagpu.so (the C ABI declared in include/mdagpu.h).

The library is the product: if it is missing or cannot be loaded every entry
point raises ``RuntimeError`` -- there is no CPU fallback in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmdagpu.so")

HOST, DEVICE = 0, 1
BOX_NONE, BOX_ORTHO, BOX_TRICLINIC = 0, 1, 2
MIC_NONE, MIC_ORTHO_ROUND, MIC_ORTHO_LAZY, MIC_TRICLINIC = 0, 1, 2, 3
ABI_VERSION = 1

_vp = C.c_void_p
_u64 = C.c_uint64
_i = C.c_int
_d = C.c_double

# name -> (restype, argtypes); mirrors include/mdagpu.h one to one
SIGNATURES = {
    "mdagpu_abi_version": (_i, []),
    "mdagpu_last_error": (C.c_char_p, []),
    "mdagpu_device_count": (_i, [C.POINTER(_i)]),
    "mdagpu_set_device": (_i, [_i]),
    "mdagpu_launch_count": (C.c_ulonglong, []),
    "mdagpu_host_to_device": (_i, [_vp, _vp, _u64, _vp]),
    "mdagpu_device_to_host": (_i, [_vp, _vp, _u64, _vp]),
    "mdagpu_profile_enable": (_i, [_i]),
    "mdagpu_profile_report": (_i, [C.c_char_p, _u64]),
    "mdagpu_distance_array": (_i, [_vp, _u64, _vp, _u64, _i, _vp, _vp, _u64, _i, _vp]),
    "mdagpu_self_distance_array": (_i, [_vp, _u64, _i, _vp, _vp, _u64, _i, _vp]),
    "mdagpu_calc_bonds": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _vp]),
    "mdagpu_calc_angles": (_i, [_vp, _vp, _vp, _u64, _i, _vp, _vp, _i, _vp]),
    "mdagpu_calc_dihedrals": (_i, [_vp, _vp, _vp, _vp, _u64, _i, _vp, _vp, _i, _vp]),
    "mdagpu_coord_transform": (_i, [_vp, _u64, _vp, _i, _vp]),
    "mdagpu_wrap": (_i, [_vp, _u64, _i, _vp, _i, _vp]),
    "mdagpu_minimize_vectors": (_i, [_vp, _u64, _i, _i, _vp, _vp, _i, _vp]),
    "mdagpu_pair_list_distances": (_i, [_vp, _u64, _vp, _u64, _u64, _vp, _vp, _u64, _i, _vp, _i, _vp, _i, _vp]),
    "mdagpu_grid_create": (_i, [_d, _vp, _u64, _vp, _vp, _i, _i, _vp, C.POINTER(_vp)]),
    "mdagpu_grid_info": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_i), C.POINTER(C.c_float)]),
    "mdagpu_grid_geometry": (_i, [_d, _vp, _vp, _i, _vp, _vp, _vp, C.POINTER(_i), C.POINTER(C.c_float)]),
    "mdagpu_grid_stencil_complete": (_i, [_d, _vp, _vp, _i, C.POINTER(_i)]),
    "mdagpu_grid_free": (None, [_vp]),
    "mdagpu_grid_search": (_i, [_vp, _vp, _u64, _i, _vp, C.POINTER(_vp)]),
    "mdagpu_grid_self_search": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "mdagpu_pairs_count": (_u64, [_vp]),
    "mdagpu_pairs_export": (_i, [_vp, _vp, _vp, _i, _vp]),
    "mdagpu_pairs_free": (None, [_vp]),
    "mdagpu_capped_bruteforce": (_i, [_vp, _u64, _vp, _u64, _d, _d, _i, _vp, _i, _vp, C.POINTER(_vp)]),
    "mdagpu_qcp": (_i, [_vp, _vp, _u64, _vp, _vp, C.POINTER(_d), _i, _vp]),
    "mdagpu_qcp_inner_product": (_i, [_vp, _vp, _u64, _vp, _vp, C.POINTER(_d), _i, _vp]),
    "mdagpu_qcp_solve": (_i, [_vp, _d, _d, _vp, C.POINTER(_d)]),
    "mdagpu_qcp_trajectory": (_i, [_vp, _u64, _u64, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "mdagpu_pairwise_rmsd": (_i, [_vp, _u64, _u64, _vp, _u64, _vp, _i, _d, _vp, _i, _vp]),
    "mdagpu_rdf_accumulate": (_i, [_vp, _u64, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _d, _d, _i, _vp, _i,
                                   _d, _d, _i, _i, _vp, _i, _vp]),
    "mdagpu_rdf_accumulate_ids": (_i, [_vp, _u64, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _d, _d, _i, _vp, _i,
                                       _d, _d, _vp, _vp, _vp, _i, _vp]),
    "mdagpu_rdf_sites_accumulate": (_i, [_vp, _u64, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _d, _d, _i, _vp, _i,
                                         _d, _d, _vp, _i, _vp]),
    "mdagpu_contacts_self": (_i, [_vp, _u64, _u64, _i, _vp, _d, _d, _vp, _vp, _i, _vp]),
    "mdagpu_frame_mean": (_i, [_vp, _u64, _u64, _vp, _u64, _vp, _i, _vp]),
    "mdagpu_group_rmsd": (_i, [_vp, _u64, _u64, _vp, _u64, _vp, _vp, _vp, _i, _vp]),
    "mdagpu_contact_fraction": (_i, [_vp, _u64, _u64, _vp, _i, _d, _d, _d, _vp, _i, _vp]),
    "mdagpu_dcd_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "mdagpu_dcd_close": (None, [_vp]),
    "mdagpu_dcd_header": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(_i), C.POINTER(C.c_int32),
                               C.POINTER(C.c_int32), C.POINTER(_d), C.POINTER(_i)]),
    "mdagpu_dcd_readframes": (_i, [_vp, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, _vp]),
}

_lib = None


class MdaGpuError(RuntimeError):
    """Raised when libmdagpu.so reports MDAGPU_ERR_CUDA / MDAGPU_ERR_NOMEM."""


def load():
    """Load libmdagpu.so (no CUDA call is made here, so this works without a GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `make -C mdanalysis_b200/csrc` "
                "(or __graft_entry__.build()); the cuda backend has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mdagpu_abi_version() != ABI_VERSION:
            raise RuntimeError("libmdagpu.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int):
    """Map a C-ABI status to the exception the reference front end would raise."""
    if rc == 0:
        return
    msg = load().mdagpu_last_error().decode("utf-8", "replace")
    if rc == -2:
        raise ValueError(msg)
    if rc == -3:
        raise MemoryError(msg)
    raise MdaGpuError(msg)


def ptr(a):
    """void* of a NumPy array (or None)."""
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)


def launch_count() -> int:
    return int(load().mdagpu_launch_count())


def profile_enable(on: bool = True):
    check(load().mdagpu_profile_enable(int(on)))


def profile_report() -> dict:
    """{kernel name: (launches, total device ms)} since the last report."""
    buf = C.create_string_buffer(1 << 16)
    check(load().mdagpu_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split("\t")
        out[name] = (int(n), float(ms))
    return out


# --- device-resident arguments -------------------------------------------------
def is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def tensor_ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def current_stream(t=None):
    """cudaStream_t of torch's current stream on the tensor's device.  Every device-resident entry
    point fetches its stream here right before the call, so this is also where the tensor's device
    is made current for the library (scratch allocations, SM count and launches follow the
    current device; a tensor on cuda:1 while cuda:0 is current would otherwise be served by the
    wrong one)."""
    import torch
    dev = t.device if t is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.index is not None and dev.index != torch.cuda.current_device():
        torch.cuda.set_device(dev)
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def same_device(*tensors):
    """ValueError unless all given tensors (None skipped) live on one CUDA device."""
    devs = {t.device for t in tensors if t is not None and is_tensor(t)}
    if len(devs) > 1:
        raise ValueError(f"all tensor arguments must be on the same device, got {sorted(map(str, devs))}")


def upload(array, device=None):
    """NumPy array -> torch CUDA tensor of the same shape / dtype through the library's staged copy
    (``mdagpu_host_to_device``): pageable arrays move at about PCIe rate instead of the driver's
    pageable path that ``torch.as_tensor(a).cuda()`` takes."""
    import numpy as np
    import torch
    a = np.ascontiguousarray(array)
    dev = torch.device("cuda", torch.cuda.current_device()) if device in (None, "cuda", True) else torch.device(device)
    out = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype, device=dev)
    if a.size:
        check(load().mdagpu_host_to_device(tensor_ptr(out), ptr(a), a.nbytes, current_stream(out)))
        torch.cuda.current_stream(dev).synchronize()      # `a` may be freed by the caller
    return out


def download(tensor, out=None):
    """torch CUDA tensor -> NumPy array through ``mdagpu_device_to_host`` (staged for pageable ``out``)."""
    import numpy as np
    import torch
    t = tensor.contiguous()
    if out is None:
        out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    if t.numel():
        check(load().mdagpu_device_to_host(ptr(out), tensor_ptr(t), out.nbytes, current_stream(t)))
    return out
