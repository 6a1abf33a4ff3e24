"""Loader of libbendy2d_b200.so (the C ABI in include/bendy2d_b200.h).

There is no CPU fallback: if the shared library is missing the import of the product fails loudly,
and if no CUDA device is present `bendy_create` fails with BENDY_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# BENDY2D_B200_LIB points at an alternative build of the same sources (tuning experiments)
LIB_PATH = os.environ.get("BENDY2D_B200_LIB") or os.path.join(_HERE, "lib", "libbendy2d_b200.so")
CSRC = os.path.join(_HERE, "csrc")

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)
u64p = C.POINTER(C.c_uint64)
intp = C.POINTER(C.c_int)

BENDY_OK = 0
ERR_NAMES = {-1: "BENDY_ERR_ARG", -2: "BENDY_ERR_LINK", -3: "BENDY_ERR_CUDA", -4: "BENDY_ERR_UNSUPPORTED",
             -5: "BENDY_ERR_NO_DEVICE"}

K_CLASSES = ["integrate", "links_local", "links_global", "links_circle", "grid_build", "narrowphase", "circles",
             "poly_prep", "poly_contact", "fused", "halo", "circle_pass"]


class ScheduleInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "n_partitions", "n_local_colours", "n_global_colours", "n_local_links", "n_global_links",
        "n_poly_partitions", "kernels_per_substep", "n_priority_partitions")]


def build(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile). Cross-compiles without a GPU."""
    srcs = [os.path.join(CSRC, f) for f in ("solver.cu", "kernels.cuh", "plan.cpp", "plan.h")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "bendy2d_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(map(os.path.getmtime, srcs))
    if stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_SIGS = None


def signatures():
    """name -> (restype, argtypes) for every symbol declared in include/bendy2d_b200.h"""
    global _SIGS
    if _SIGS is not None:
        return _SIGS
    vp, sz, fl, i, u32, u16 = C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_uint32, C.c_uint16
    _SIGS = {
        "bendy_create": (vp, [i]),
        "bendy_destroy": (None, [vp]),
        "bendy_clone": (vp, [vp]),
        "bendy_last_error": (C.c_char_p, [vp]),
        "bendy_abi_version": (i, []),
        "bendy_save_snapshot": (i, [vp, C.c_char_p]),
        "bendy_load_snapshot": (vp, [C.c_char_p, i]),
        "bendy_get_last_update_args": (i, [vp, f32p, intp]),
        "bendy_add_particles": (i, [vp, f32p, sz]),
        "bendy_add_circles": (i, [vp, f32p, f32p, f32p, f32p, sz]),
        "bendy_add_polygon": (i, [vp, f32p, f32p, f32p, sz, u32p, f32p, sz, i, fl, fl]),
        "bendy_add_particle_links": (i, [vp, u32p, f32p, sz]),
        "bendy_add_circle_links": (i, [vp, u32p, f32p, sz]),
        "bendy_update": (i, [vp, fl, fl, fl, fl, fl, fl, fl]),
        "bendy_update_n": (i, [vp, u32, fl, fl, fl, fl, fl, fl, fl]),
        "bendy_synchronize": (i, [vp]),
        "bendy_particle_len": (sz, [vp]),
        "bendy_circle_len": (sz, [vp]),
        "bendy_polygon_len": (sz, [vp]),
        "bendy_particle_link_len": (sz, [vp]),
        "bendy_circle_link_len": (sz, [vp]),
        "bendy_polygon_point_len": (sz, [vp, sz]),
        "bendy_polygon_link_len": (sz, [vp, sz]),
        "bendy_read_particles": (i, [vp, sz, sz, f32p, f32p]),
        "bendy_read_circles": (i, [vp, sz, sz, f32p, f32p, f32p]),
        "bendy_read_polygon": (i, [vp, sz, f32p, f32p, f32p, intp]),
        "bendy_read_particle_links": (i, [vp, sz, sz, u32p, f32p]),
        "bendy_read_circle_links": (i, [vp, sz, sz, u32p, f32p]),
        "bendy_read_polygon_links": (i, [vp, sz, u32p, f32p]),
        "bendy_set_sub_steps": (i, [vp, u16]),
        "bendy_write_particles": (i, [vp, sz, sz, f32p, f32p]),
        "bendy_set_particle_radius": (i, [vp, fl]),
        "bendy_set_grid_cell": (i, [vp, fl]),
        "bendy_set_polygon_contact": (i, [vp, i]),
        "bendy_set_particle_inv_mass": (i, [vp, sz, sz, f32p]),
        "bendy_set_circle_inv_mass": (i, [vp, sz, sz, f32p]),
        "bendy_set_plan_params": (i, [vp, u32, u32]),
        "bendy_set_link_schedule": (i, [vp, i]),
        "bendy_get_schedule_info": (i, [vp, C.POINTER(ScheduleInfo)]),
        "bendy_get_link_order": (i, [vp, u32p, sz]),
        "bendy_get_point_rank": (i, [vp, u32p, sz]),
        "bendy_get_grid": (i, [vp, fl, fl, fl, fl, f32p, f32p, f32p, intp, intp]),
        "bendy_set_profiling": (i, [vp, i]),
        "bendy_get_kernel_times": (i, [vp, f64p, u64p, i, i]),
        "bendy_get_stats": (i, [vp, u64p, i]),
        "bendy_launch_count": (C.c_uint64, [vp]),
        "bendy_timer_start": (i, [vp]),
        "bendy_timer_stop": (i, [vp, f32p]),
        "bendy_get_stream": (vp, [vp]),
        "bendy_get_device": (i, [vp]),
        "bendy_get_device_buffers": (i, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]),
        "bendy_halo_configure": (i, [vp, u32, fl, fl, fl, fl]),
        "bendy_set_grid_window": (i, [vp, fl, fl]),
        "bendy_strip_set_cross_links": (i, [vp, sz, u32p, u32p, C.POINTER(C.c_uint8), f32p, u32, u32p, sz, u32p, sz, u32p, sz, sz]),
        "bendy_nccl_unique_id": (i, [vp]),
        "bendy_halo_comm_nccl": (i, [vp, vp, i, i]),
        "bendy_halo_connect_local": (i, [vp, vp]),
        "bendy_update_group": (i, [C.POINTER(vp), i, u32, fl, fl, fl, fl, fl, fl, fl]),
        "bendy_halo_stats": (i, [vp, u32p, u32p, u32p, u32p]),
        "bendy_plan_links": (i, [sz, u32p, sz, u32, u32, u32p, u32p, u32p, u32p, C.POINTER(ScheduleInfo)]),
        "bendy_plan_links_scheduled": (i, [sz, u32p, sz, u32, u32, i, u32p, u32p, u32p, u32p, C.POINTER(ScheduleInfo)]),
        "bendy_debug_normalize": (i, [i, f32p, f32p, f32p, sz, f32p, f32p]),
    }
    return _SIGS


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(bendy2d_b200 has no CPU fallback)")
    # tests/cuemu builds the same sources against a CPU emulation of CUDA to check kernel logic without a GPU.
    # That library is test infrastructure: it is only accepted when the test harness says so explicitly, so
    # that no configuration mistake can turn it into a CPU path of the product.
    if os.path.basename(LIB_PATH).endswith("_emu.so") and os.environ.get("BENDY_CUDA_EMU") != "1":
        raise ImportError(f"{LIB_PATH} is the CPU emulation build used by the tests, not a product library "
                          "(bendy2d_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in signatures().items():
        fn = getattr(L, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
