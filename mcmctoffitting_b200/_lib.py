"""ctypes binding of libtofgpu.so (include/tofgpu.h).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

MAX_DIM, MAX_RUNS, MAX_MATERIALS = 16, 8, 8
ABI_VERSION = 3

_dp = C.POINTER(C.c_double)


class TofConfig(C.Structure):
    """Mirror of ``struct tof_config`` (include/tofgpu.h); size checked against the library."""
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("device", C.c_int32), ("ode_mode", C.c_int32),
        ("ode_substeps", C.c_int32), ("ode_from_zero", C.c_int32), ("prior_strict", C.c_int32),
        ("nan_to_neginf", C.c_int32), ("ndim", C.c_int32), ("n_runs", C.c_int32), ("x_bins", C.c_int32),
        ("e_bins", C.c_int32), ("n_taps", C.c_int32), ("n_zero_deg", C.c_int32), ("n_materials", C.c_int32),
        ("n_xs", C.c_int32),
        ("n_samples", C.c_int64), ("n_ev_per_loop", C.c_int64), ("n_loops", C.c_int64),
        ("x_min", C.c_double), ("x_max", C.c_double), ("e_min", C.c_double), ("e_max", C.c_double),
        ("speed_of_light", C.c_double), ("mass_deuteron", C.c_double), ("mass_neutron", C.c_double),
        ("mass_he3", C.c_double), ("q_ddn", C.c_double), ("cell_length", C.c_double),
        ("simple_neutron_base", C.c_double),
        ("bethe_A", C.c_double * MAX_MATERIALS), ("bethe_B", C.c_double * MAX_MATERIALS),
        ("prior_lo", C.c_double * MAX_DIM), ("prior_hi", C.c_double * MAX_DIM),
        ("tof_bins", C.c_int32 * MAX_RUNS), ("tof_min", C.c_double * MAX_RUNS), ("tof_max", C.c_double * MAX_RUNS),
        ("x_centers", _dp), ("e_centers", _dp), ("neutron_speed", _dp), ("neutron_dist", _dp),
        ("xs_breaks", _dp), ("xs_coefs", _dp), ("taps", _dp), ("zero_deg_times", _dp), ("zero_deg_weights", _dp),
        ("t1_q", C.c_int32), ("t1_key_lo", C.c_int32), ("t1_n", C.c_int32), ("rng_degree", C.c_int32),
        ("rng_n", C.c_int32), ("rng_lut_n", C.c_int32),
        ("rng_sign", C.c_double), ("rng_u_max", C.c_double), ("e_tab_lo", C.c_double), ("e_tab_hi", C.c_double),
        ("t1_coefs", _dp), ("rng_breaks", _dp), ("rng_bins", C.POINTER(C.c_int32)), ("rng_coefs", _dp),
        ("rng_lut", C.POINTER(C.c_uint16)),
        ("stop_n", C.c_int32), ("n_taps2", C.c_int32), ("stop_lo", C.c_double), ("stop_step", C.c_double),
        ("beam_energy", C.c_double), ("stop_coefs", _dp), ("attenuation", _dp), ("taps2", _dp),
        ("precision", C.c_int32),
    ]


class TofStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("evaluations", C.c_int64), ("nan_results", C.c_int64),
                ("sm_count", C.c_int32), ("smem_bytes", C.c_int32), ("threads", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("band_ctas_per_sm", C.c_int32), ("band_cells", C.c_int32),
                ("band_queued_last", C.c_int64), ("fp32_active", C.c_int32), ("model_launches_per_call", C.c_int32),
                ("wide_last", C.c_int64)]


class TofError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libtofgpu error %d: %s" % (code, msg))
        self.code = code


_LIB = None

# every symbol include/tofgpu.h declares: (name, restype, argtypes)
_vp = C.c_void_p
SIGNATURES = {
    "tof_abi_version": (C.c_int, []),
    "tof_sizeof_config": (C.c_int, []),
    "tof_create": (C.c_int, [C.POINTER(TofConfig), C.POINTER(_vp)]),
    "tof_destroy": (None, [_vp]),
    "tof_last_error": (C.c_char_p, [_vp]),
    "tof_set_observables": (C.c_int, [_vp, C.c_int, _dp, C.c_int]),
    "tof_set_draws": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int64]),
    "tof_set_draw_mode": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_uint64]),
    "tof_generate_draws": (C.c_int, [_vp, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, _dp, C.c_int64]),
    "tof_lnprob_batch": (C.c_int, [_vp, _dp, C.c_int64, _dp]),
    "tof_lnprob_batch_device": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp]),
    "tof_model_batch": (C.c_int, [_vp, _dp, C.c_int64, C.c_int, C.c_int, _dp]),
    "tof_cell_counts_batch": (C.c_int, [_vp, _dp, C.c_int64, C.c_int, C.POINTER(C.c_int64)]),
    "tof_deuteron_counts_batch": (C.c_int, [_vp, _dp, C.c_int64, C.c_int, C.POINTER(C.c_int64)]),
    "tof_stretch_propose": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_double, C.c_uint64,
                                      C.c_int64, C.c_int, _vp, _vp, _vp]),
    "tof_stretch_accept": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_uint64, C.c_int64,
                                     C.c_int, _vp, _vp]),
    "tof_ensemble_step": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_double, C.c_uint64, C.c_int64, _vp, _vp]),
    "tof_ensemble_half_step": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_uint64, C.c_int64,
                                         _vp, _vp]),
    "tof_get_stats": (C.c_int, [_vp, C.POINTER(TofStats)]),
    "tof_set_timing": (C.c_int, [_vp, C.c_int]),
    "tof_set_stage_timing": (C.c_int, [_vp, C.c_int]),
    "tof_get_stage_cycles": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "tof_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "tof_measure_fp64_peak": (C.c_int, [_vp, C.POINTER(C.c_double)]),
}


def header_symbols() -> list[str]:
    """Function names declared in include/tofgpu.h (used by the export test)."""
    hdr = os.path.join(os.path.dirname(_build.PKG_DIR), "include", "tofgpu.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tof_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libtofgpu.so (building it with nvcc when it is missing).  Raises when neither works:
    there is no other implementation to fall back to."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    override = os.environ.get("TOFGPU_LIB")            # e.g. the checked build (build.py --checked); same ABI, same kernels
    if override:
        if not os.path.exists(override):
            raise RuntimeError("TOFGPU_LIB=%s does not exist" % override)
        path = override
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError("libtofgpu.so is not built; run `python -m mcmctoffitting_b200.build`")
        _build.build_library()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.tof_abi_version() != ABI_VERSION:
        raise RuntimeError("libtofgpu.so ABI %d != binding ABI %d" % (lib.tof_abi_version(), ABI_VERSION))
    if lib.tof_sizeof_config() != C.sizeof(TofConfig):
        raise RuntimeError("tof_config layout mismatch: library %d bytes, binding %d bytes" %
                           (lib.tof_sizeof_config(), C.sizeof(TofConfig)))
    _LIB = lib
    return lib


def check(lib, ctx, rc: int) -> None:
    if rc != 0:
        msg = lib.tof_last_error(ctx)
        raise TofError(rc, msg.decode() if msg else "")
