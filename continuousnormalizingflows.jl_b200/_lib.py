"""ctypes binding of ``libicnf_b200.so`` (the C ABI in ``include/icnf_b200.h``).

This is the stand-in for the Julia ``ccall`` glue (``julia/B200Mode.jl``): every
call the host mirror makes goes through the same exported symbols a Julia caller
would bind.  There is no fallback of any kind: if the shared library is missing
or cannot be loaded, importing this module raises.
"""

from __future__ import annotations

import ctypes as C
import os

ICNF_MAX_LAYERS = 8
ICNF_ABI_VERSION = 1

# icnf_status
OK, ERR_INVALID, ERR_CUDA, ERR_MAX_STEPS, ERR_DT_UNDERFLOW, ERR_NONFINITE, ERR_NO_DEVICE, ERR_UNSUPPORTED = range(8)
# icnf_mode
MODE_TEST, MODE_TRAIN_REG, MODE_TRAIN_NOREG = 0, 1, 2
# icnf_activation
ACT = {"softplus": 0, "tanh": 1, "sigmoid": 2, "identity": 3}
# icnf_eps_kind
EPS_SUPPLIED, EPS_GAUSSIAN, EPS_RADEMACHER = 0, 1, 2
EPS = {"supplied": 0, "gaussian": 1, "rademacher": 2}
# icnf_precision
PRECISION = {"fp32": 0, "bf16_tc": 1, "bf16x3_tc": 2}


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("nvars", C.c_int32), ("naug", C.c_int32), ("ncond", C.c_int32),
        ("autonomous", C.c_int32), ("n_layers", C.c_int32), ("sizes", C.c_int32 * (ICNF_MAX_LAYERS + 1)),
        ("activation", C.c_int32), ("lambda1", C.c_float), ("lambda2", C.c_float), ("lambda3", C.c_float),
        ("reg_squared", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
    ]


class Solver(C.Structure):
    _fields_ = [
        ("adaptive", C.c_int32), ("dt", C.c_float), ("reltol", C.c_float), ("abstol", C.c_float),
        ("max_steps", C.c_int32), ("beta1", C.c_float), ("beta2", C.c_float), ("gamma", C.c_float),
        ("qmin", C.c_float), ("qmax", C.c_float), ("qsteady_min", C.c_float), ("qsteady_max", C.c_float),
        ("qoldinit", C.c_float), ("alg", C.c_int32),
    ]


class Noise(C.Structure):
    _fields_ = [("kind", C.c_int32), ("seed", C.c_uint64), ("sample_offset", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("naccept", C.c_int32), ("nreject", C.c_int32), ("nf", C.c_int32), ("status", C.c_int32),
                ("t_final", C.c_float), ("dt_last", C.c_float)]


LIB_NAME = "libicnf_b200.so"
# ICNF_B200_LIB points at an alternative build of the same library (tuning sweeps)
LIB_PATH = os.environ.get("ICNF_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

# every symbol include/icnf_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_F = C.c_void_p  # float* (host or device address)
SYMBOLS = [
    ("icnf_version", C.c_char_p, []),
    ("icnf_device_count", C.c_int, []),
    ("icnf_create", C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    ("icnf_destroy", None, [_P]),
    ("icnf_last_error", C.c_char_p, [_P]),
    ("icnf_n_params", C.c_int64, [_P]),
    ("icnf_n_state", C.c_int32, [_P]),
    ("icnf_kernel_family", C.c_char_p, [_P]),
    ("icnf_solve_path", C.c_char_p, [_P, C.c_int]),
    ("icnf_set_params", C.c_int, [_P, _F, C.c_int64]),
    ("icnf_set_params_dev", C.c_int, [_P, _F, C.c_int64, _P]),
    ("icnf_rhs", C.c_int, [_P, C.c_int, C.c_float, _F, _F, _F, _F, C.c_int64]),
    ("icnf_rhs_dev", C.c_int, [_P, C.c_int, C.c_float, _F, _F, _F, _F, C.c_int64, _P]),
    ("icnf_solve", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F, _F,
                             C.POINTER(Stats), C.c_int64]),
    ("icnf_solve_dev", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F, _F,
                                 _P, C.c_int64, _P]),
    ("icnf_inference", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F, _F,
                                 _F, C.POINTER(Stats), C.c_int64]),
    ("icnf_inference_dev", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                                     _F, _F, _P, C.c_int64, _P]),
    ("icnf_generate", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F, _F,
                                C.POINTER(Stats), C.c_int64]),
    ("icnf_generate_dev", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                                    _F, _P, C.c_int64, _P]),
    ("icnf_loss", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                            C.POINTER(C.c_float), C.POINTER(Stats), C.c_int64, C.c_int64]),
    ("icnf_loss_grad", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                                 C.POINTER(C.c_float), _F, _F, C.POINTER(Stats), C.c_int64, C.c_int64]),
    ("icnf_loss_grad_dev", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                                     _F, _F, _F, _P, C.c_int64, C.c_int64, _P]),
    ("icnf_steer_tspan", C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64, C.POINTER(C.c_float)]),
    ("icnf_group_unique_id", C.c_int, [_P]),
    ("icnf_group_join", C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    ("icnf_create_group", C.c_int, [C.POINTER(_P), C.c_int32]),
    ("icnf_group_leave", C.c_int, [_P]),
    ("icnf_group_info", C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("icnf_group_set_global_norm", C.c_int, [_P, C.c_int]),
    ("icnf_group_start", C.c_int, []),
    ("icnf_group_end", C.c_int, []),
    ("icnf_loss_grad_dp", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                                    C.POINTER(C.c_float), _F, _F, C.POINTER(Stats), C.c_int64, C.c_int64]),
    ("icnf_loss_grad_dp_dev", C.c_int, [_P, C.c_int, C.POINTER(Solver), C.c_float, C.c_float, _F, C.POINTER(Noise), _F, _F,
                                        _F, _F, _F, _P, C.c_int64, C.c_int64, _P]),
    ("icnf_launch_count", C.c_int64, [_P]),
    ("icnf_set_profiling", C.c_int, [_P, C.c_int]),
    ("icnf_kernel_times", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("icnf_measure_fp32_peak", C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    ("icnf_adam_step_dev", C.c_int, [_F, _F, _F, _F, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_float, _P]),
    ("icnf_tc_gemm_selftest", C.c_int, [C.c_int, C.c_int, C.c_int, _F, _F, _F, C.c_int]),
    ("icnf_tc_wgrad_selftest", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _F, _F, _F, _F, _F, C.c_int, C.c_int]),
    ("icnf_backward_plan", C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
]


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(
            f"{LIB_NAME} not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C continuousnormalizingflows.jl_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = load()


class ICNFError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"icnf_b200 error {code}: {message}")
        self.code = code
