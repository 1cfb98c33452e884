"""ctypes binding of libbfilt.so (include/bfilt.h).  No CPU fallback: a missing library is an error."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFILT_LIB", os.path.join(_HERE, "libbfilt.so"))

BF_OK, BF_EINVAL, BF_ECUDA, BF_ENOMODEL, BF_ENOMEM, BF_ECOMM = 0, -1, -2, -3, -4, -5
MODEL_UNGM, MODEL_BOT, MODEL_L63, MODEL_L96, MODEL_LG, MODEL_SIN1D, MODEL_MSV = 1, 2, 3, 4, 5, 6, 7
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC = 0, 1
PATH_AUTO, PATH_RESIDENT, PATH_STREAMED = 0, 1, 2
GSF_PREDICT_FIRST, GSF_NO_JITTER, GSF_COV_PLUS_P = 1, 2, 4
LEGACY_UKF, LEGACY_MCF = 1, 2
SHARD_BLOB_BYTES, NCCL_ID_BYTES = 512, 128

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("dx", C.c_int32), ("dq", C.c_int32), ("dy", C.c_int32), ("dr", C.c_int32),
                ("n_theta", C.c_int32), ("theta", _fp), ("q0", _fp), ("Q", _fp), ("r0", _fp), ("R", _fp),
                ("m0", _fp), ("P0", _fp)]


class PfConfig(C.Structure):
    _fields_ = [("num_particles", C.c_int64), ("num_trials", C.c_int32), ("ess_threshold", C.c_float),
                ("resampler", C.c_int32), ("path", C.c_int32), ("seed", C.c_uint64), ("device", C.c_int32),
                ("trial_offset", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32)]


class PfIO(C.Structure):
    _fields_ = [("T", C.c_int32), ("t_offset", C.c_int32), ("resume", C.c_int32), ("emissions_per_trial", C.c_int32),
                ("emissions", C.c_void_p), ("inputs", C.c_void_p), ("init_particles", C.c_void_p),
                ("init_weights", C.c_void_p), ("noise", C.c_void_p), ("resample_u", C.c_void_p),
                ("out_particles", C.c_void_p), ("out_weights", C.c_void_p), ("out_propagated", C.c_void_p),
                ("out_norm_weights", C.c_void_p), ("out_ancestors", C.c_void_p), ("out_mean", C.c_void_p),
                ("out_loglik", C.c_void_p), ("out_ess", C.c_void_p), ("out_resampled", C.c_void_p),
                ("final_particles", C.c_void_p), ("final_weights", C.c_void_p)]


class GsfConfig(C.Structure):
    _fields_ = [("num_components", C.c_int32), ("num_trials", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32)]


class GsfIO(C.Structure):
    _fields_ = [("T", C.c_int32), ("resume", C.c_int32), ("emissions_per_trial", C.c_int32), ("reserved", C.c_int32),
                ("emissions", C.c_void_p), ("inputs", C.c_void_p), ("init_means", C.c_void_p), ("init_covs", C.c_void_p),
                ("init_weights", C.c_void_p),
                ("out_weights", C.c_void_p), ("out_means", C.c_void_p), ("out_covs", C.c_void_p),
                ("out_pred_means", C.c_void_p), ("out_pred_covs", C.c_void_p), ("out_loglik", C.c_void_p),
                ("out_mix_mean", C.c_void_p), ("Q_t", C.c_void_p), ("R_t", C.c_void_p), ("mc_draws", C.c_void_p)]


class AgsfConfig(C.Structure):
    _fields_ = [("num_components", C.c_int32), ("num_branch1", C.c_int32), ("num_branch2", C.c_int32), ("num_trials", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32), ("alpha1", C.c_float), ("alpha2", C.c_float), ("seed", C.c_uint64)]


class AgsfIO(C.Structure):
    _fields_ = [("T", C.c_int32), ("emissions_per_trial", C.c_int32), ("rand_per_step", C.c_int32), ("reserved", C.c_int32),
                ("emissions", C.c_void_p), ("inputs", C.c_void_p), ("init_means", C.c_void_p), ("init_covs", C.c_void_p),
                ("z_iid", C.c_void_p), ("s_iid", C.c_void_p), ("res_u", C.c_void_p),
                ("out_weights", C.c_void_p), ("out_means", C.c_void_p), ("out_covs", C.c_void_p), ("out_loglik", C.c_void_p),
                ("out_pre_weights", C.c_void_p), ("out_updated_means", C.c_void_p), ("out_ancestors", C.c_void_p)]


# every symbol include/bfilt.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("bf_pf_create", C.c_int, [C.POINTER(ModelDesc), C.POINTER(PfConfig), C.POINTER(C.c_void_p)]),
    ("bf_pf_run", C.c_int, [C.c_void_p, C.POINTER(PfIO), C.c_void_p]),
    ("bf_pf_step", C.c_int, [C.c_void_p, C.POINTER(PfIO), C.c_void_p]),
    ("bf_pf_run_host", C.c_int, [C.c_void_p, C.POINTER(PfIO)]),
    ("bf_pf_kernel_launches", C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("bf_pf_phase_times", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int32]),
    ("bf_pf_destroy", C.c_int, [C.c_void_p]),
    ("bf_pf_shard_export", C.c_int, [C.c_void_p, C.c_void_p]),
    ("bf_pf_shard_import", C.c_int, [C.c_void_p, C.c_void_p]),
    ("bf_nccl_unique_id", C.c_int, [C.c_void_p]),
    ("bf_pf_shard_comm_init", C.c_int, [C.c_void_p, C.c_void_p]),
    ("bf_pf_shard_comm_ready", C.c_int, [C.c_void_p]),
    ("bf_gsf_create", C.c_int, [C.POINTER(ModelDesc), C.POINTER(GsfConfig), C.POINTER(C.c_void_p)]),
    ("bf_gsf_run", C.c_int, [C.c_void_p, C.POINTER(GsfIO), C.c_void_p]),
    ("bf_gsf_step", C.c_int, [C.c_void_p, C.POINTER(GsfIO), C.c_void_p]),
    ("bf_gsf_run_host", C.c_int, [C.c_void_p, C.POINTER(GsfIO)]),
    ("bf_gsf_set_unscented", C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    ("bf_gsf_set_legacy_moments", C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_uint64]),
    ("bf_gsf_kernel_launches", C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("bf_gsf_destroy", C.c_int, [C.c_void_p]),
    ("bf_agsf_create", C.c_int, [C.POINTER(ModelDesc), C.POINTER(AgsfConfig), C.POINTER(C.c_void_p)]),
    ("bf_agsf_run", C.c_int, [C.c_void_p, C.POINTER(AgsfIO), C.c_void_p]),
    ("bf_agsf_run_host", C.c_int, [C.c_void_p, C.POINTER(AgsfIO)]),
    ("bf_agsf_destroy", C.c_int, [C.c_void_p]),
    ("bf_optimal_resampling", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bf_cumsum", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("bf_resample", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    ("bf_model_eval", C.c_int, [C.POINTER(ModelDesc), C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_float,
                                C.c_void_p, C.c_void_p]),
    ("bf_ssm_sample", C.c_int, [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bf_measure_fp32_peak", C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    ("bf_trim_memory", C.c_int, [C.c_int32]),
    ("bf_last_error", C.c_char_p, []),
    ("bf_abi_version", C.c_int, []),
]

_lib = None


class BfiltError(RuntimeError):
    pass


def load():
    """Load libbfilt.so (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BfiltError(
            f"{LIB_PATH} is missing: build it with bayesianfiltering_b200/csrc/build.sh "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == BF_OK:
        return
    msg = load().bf_last_error().decode("utf-8", "replace")
    if rc in (BF_EINVAL, BF_ENOMODEL):
        raise ValueError(f"libbfilt: {msg}")
    if rc == BF_ENOMEM:
        raise MemoryError(f"libbfilt: {msg}")
    raise BfiltError(f"libbfilt ({rc}): {msg}")


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


class HostModel:
    """Owns the fp32 host arrays a bf_model_desc points at."""

    def __init__(self, model_id, dx, dq, dy, dr, theta, q0, Q, r0, R, m0, P0):
        self.model_id, self.dx, self.dq, self.dy, self.dr = int(model_id), int(dx), int(dq), int(dy), int(dr)
        self.theta = _f32(theta).ravel()
        self.q0 = _f32(q0).reshape(-1)
        self.Q = _f32(Q).reshape(-1)
        self.r0 = _f32(r0).reshape(-1)
        self.R = _f32(R).reshape(-1)
        self.m0 = _f32(m0).reshape(-1)
        self.P0 = _f32(P0).reshape(-1)
        exp = dict(q0=dq, Q=dq * dq, r0=dr, R=dr * dr, m0=dx, P0=dx * dx)
        for k, n in exp.items():
            if getattr(self, k).size != n:
                raise ValueError(f"model array {k} has {getattr(self, k).size} elements, expected {n}")
        p = lambda a: a.ctypes.data_as(_fp)
        self.desc = ModelDesc(self.model_id, self.dx, self.dq, self.dy, self.dr, self.theta.size, p(self.theta),
                              p(self.q0), p(self.Q), p(self.r0), p(self.R), p(self.m0), p(self.P0))
