"""ctypes binding of libmopoe_b200.so (include/mopoe_b200.h).  No CPU fallback: a missing library or
a missing CUDA device is a hard error on every compute call."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOPOE_LIB_PATH: load another build of the library (profiling builds made with `make EXTRA=-D... OUT=...`)
LIB_PATH = os.environ.get("MOPOE_LIB_PATH") or os.path.join(_HERE, "libmopoe_b200.so")

MAX_MODS = 4
MAX_SUBSETS = 15
HIDDEN = 256
N_SCALARS = 64
METHODS = {"poe": 0, "moe": 1, "joint_elbo": 2, "jsd": 3}
LIKELIHOODS = {"normal": 0, "laplace": 1}

# mopoe_scalar_index
S_TOTAL_LOSS, S_JOINT_DIV, S_NLL, S_NLL_UNI, S_KLD_SUBSET, S_KLD_STYLE, S_MEAN_HEAD = 0, 1, 2, 6, 10, 25, 29
S_N_ROWS, S_PRESENT, S_JSD_DIV = 45, 46, 47
STREAM_DAA_BASE, STREAM_DAA_SCORE, STREAM_DAA_AVATAR, STREAM_TRAIN, STREAM_FORWARD = 1, 2, 3, 4, 5


class ModelDesc(C.Structure):
    _fields_ = [("n_mods", C.c_int32), ("dims", C.c_int32 * MAX_MODS), ("style_dims", C.c_int32 * MAX_MODS),
                ("latent_dim", C.c_int32), ("hidden", C.c_int32), ("n_hidden_enc", C.c_int32),
                ("n_hidden_dec", C.c_int32), ("method", C.c_int32), ("likelihood", C.c_int32),
                ("scale_mode", C.c_int32), ("learn_output_scale", C.c_int32),
                ("name_rank", C.c_int32 * MAX_MODS), ("beta", C.c_float), ("beta_style", C.c_float),
                ("beta_content", C.c_float)]


MAX_LAYERS = 4


class ParamLayout(C.Structure):
    _fields_ = ([(n, C.c_int64 * MAX_MODS) for n in
                 ("enc_w1", "enc_b1", "enc_wh", "enc_bh", "dec_w", "dec_b", "dec_lv")] + [("total", C.c_int64)] +
                [(n, (C.c_int64 * (MAX_LAYERS - 1)) * MAX_MODS) for n in ("enc_wx", "enc_bx")] +
                [(n, (C.c_int64 * MAX_LAYERS) * MAX_MODS) for n in ("dec_hw", "dec_hb")] +
                [(n, C.c_int64 * MAX_MODS) for n in ("dec_lvw", "dec_lvb")])


class BatchDesc(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("present_mask", C.c_int32), ("n_mix", C.c_int32),
                ("joint_bounds", C.c_int32 * (MAX_SUBSETS + 1)),
                ("moe_bounds", (C.c_int32 * (MAX_MODS + 1)) * (MAX_MODS + 1)), ("row_offset", C.c_int64),
                ("owner_div", C.c_int32), ("owner_mod", C.c_int32)]


class ForwardOut(C.Structure):
    _fields_ = [("enc_heads", C.c_void_p * MAX_MODS), ("subset_mu", C.c_void_p), ("subset_logvar", C.c_void_p),
                ("joint_mu", C.c_void_p), ("joint_logvar", C.c_void_p), ("z", C.c_void_p),
                ("z_style", C.c_void_p * MAX_MODS), ("rec_loc", C.c_void_p * MAX_MODS), ("scalars", C.c_void_p),
                ("rec_logvar", C.c_void_p * MAX_MODS)]


class DaaDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_val", "val_begin", "n_val_total", "n_subjects", "n_samples", "n_base",
                                         "src_mod", "dst_mod", "sample_latents", "reg_method", "base_mode",
                                         "unit_begin", "unit_end", "score_mode")]


MAX_PEERS = 8


class TableExchangeDesc(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("root", C.c_int32), ("reserved", C.c_int32),
                ("elems_local", C.c_int64), ("elem_offset", C.c_int64),
                ("elems_total", C.c_int64), ("peer_base", C.c_void_p * MAX_PEERS)]


class MopoeError(RuntimeError):
    pass


_lib = None


def build(verbose=False):
    """Compile libmopoe_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise MopoeError("building libmopoe_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MopoeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback for the MoPoE hot path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    L.mopoe_last_error.restype = C.c_char_p
    L.mopoe_version.restype = C.c_int
    L.mopoe_device_count.restype = C.c_int
    L.mopoe_param_layout_of.argtypes = [C.POINTER(ModelDesc), C.POINTER(ParamLayout)]
    L.mopoe_workspace_bytes.argtypes = [C.POINTER(ModelDesc), i64]
    L.mopoe_workspace_bytes.restype = i64
    L.mopoe_forward.argtypes = [C.POINTER(ModelDesc), vp, C.POINTER(BatchDesc), C.POINTER(vp), vp, u64, C.c_int,
                                C.c_int, C.c_int, C.POINTER(ForwardOut), vp, i64, vp]
    L.mopoe_train_steps.argtypes = [C.POINTER(ModelDesc), vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, i32,
                                    i64, vp, u64, C.c_int, f32, f32, f32, f32, vp, C.POINTER(ForwardOut), vp, i64, vp]
    L.mopoe_daa_workspace_bytes.argtypes = [C.POINTER(ModelDesc), C.POINTER(DaaDesc)]
    L.mopoe_daa_workspace_bytes.restype = i64
    L.mopoe_daa_sweep.argtypes = [C.POINTER(ModelDesc), vp, C.POINTER(DaaDesc), C.POINTER(BatchDesc), C.POINTER(vp),
                                  vp, vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.mopoe_daa_regression.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.mopoe_philox_normal.argtypes = [u64, u64, i64, i64, vp, vp]
    L.mopoe_daa_last_impl.restype = C.c_int
    L.mopoe_table_exchange_bytes.argtypes = [i64]
    L.mopoe_table_exchange_bytes.restype = i64
    L.mopoe_daa_exchange_tables.argtypes = [C.POINTER(TableExchangeDesc), vp, vp, vp]
    L.mopoe_daa_exchange_tables.restype = C.c_int
    L.mopoe_train_last_impl.restype = C.c_int
    L.mopoe_daa_read_phases.argtypes = [C.POINTER(ModelDesc), C.POINTER(DaaDesc), vp, C.POINTER(C.c_int64)]
    L.mopoe_daa_read_phases.restype = C.c_int
    L.mopoe_daa_status.argtypes = [C.POINTER(ModelDesc), C.POINTER(DaaDesc), vp, vp]
    L.mopoe_daa_status.restype = C.c_int
    L.mopoe_profile_enable.argtypes = [C.c_int]
    L.mopoe_profile_enable.restype = C.c_int
    L.mopoe_daa_last_kernel_ms.argtypes = [C.POINTER(C.c_float)]
    L.mopoe_daa_last_kernel_ms.restype = C.c_int
    L.mopoe_rsa_cmat.argtypes = [i32, i32, vp, i32, vp, vp]
    L.mopoe_rsa_cmat.restype = C.c_int
    L.mopoe_rsa_kendall_workspace_bytes.argtypes = [i32, i32]
    L.mopoe_rsa_kendall_workspace_bytes.restype = i64
    L.mopoe_rsa_kendall.argtypes = [i32, i32, vp, vp, vp, vp, i64, vp]
    L.mopoe_rsa_kendall.restype = C.c_int
    for fn in ("mopoe_param_layout_of", "mopoe_forward", "mopoe_train_steps", "mopoe_daa_sweep",
               "mopoe_daa_regression", "mopoe_philox_normal"):
        getattr(L, fn).restype = C.c_int
    _lib = L
    return L


SELFTEST_LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), "libmopoe_b200_selftest.so")
_selftest = None


def selftest_lib():
    """libmopoe_b200_selftest.so: the tcgen05 building-block self-test (test-only, not part of the product library)."""
    global _selftest
    if _selftest is None:
        if not os.path.isfile(SELFTEST_LIB_PATH):
            raise MopoeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % SELFTEST_LIB_PATH)
        T = C.CDLL(SELFTEST_LIB_PATH)
        T.mopoe_umma_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        T.mopoe_umma_selftest.restype = C.c_int
        T.mopoe_last_error.restype = C.c_char_p
        _selftest = T
    return _selftest


def check(rc):
    if rc != 0:
        raise MopoeError("libmopoe_b200: error %d: %s" % (rc, lib().mopoe_last_error().decode()))


EXPORTED = ["mopoe_last_error", "mopoe_version", "mopoe_device_count", "mopoe_param_layout_of",
            "mopoe_workspace_bytes", "mopoe_forward", "mopoe_train_steps", "mopoe_daa_workspace_bytes",
            "mopoe_daa_sweep", "mopoe_daa_regression", "mopoe_philox_normal", "mopoe_profile_enable",
            "mopoe_daa_last_kernel_ms", "mopoe_daa_last_impl", "mopoe_daa_read_phases",
            "mopoe_daa_status", "mopoe_train_last_impl", "mopoe_table_exchange_bytes", "mopoe_daa_exchange_tables",
            "mopoe_rsa_cmat", "mopoe_rsa_kendall_workspace_bytes", "mopoe_rsa_kendall"]
