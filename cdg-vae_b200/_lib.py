"""ctypes binding of include/cdgvae.h.  Loading fails loudly: there is no CPU or PyTorch fallback."""
import ctypes as C
import os

from . import build as _build

MAX_NODE, MAX_DEC, MAX_FLOW, MAX_LAYERS, MAX_SPANS, MAX_SEG = 8, 8, 2, 4, 64, 32
SCM = {"linear": 0, "nonlinear": 1}
GEMM_MODES = {"auto": 0, "simt": 1, "tc3x": 2, "tc1x": 3, "bf3x": 4}
TAB_KIND = {"loan": 0, "adult": 1, "covtype": 2, "tvae": 3}
PROF_CATS = ["enc0_fwd", "dec2_fwd", "dec2_dgrad", "dec2_wgrad", "enc0_wgrad", "gemm_other", "latent", "recon", "misc"]


class Linear(C.Structure):
    _fields_ = [("w", C.c_int64), ("b", C.c_int64), ("in_", C.c_int32), ("out", C.c_int32)]


class AdamArgs(C.Structure):
    _fields_ = [("n_seg", C.c_int32), ("seg_off", C.c_int64 * MAX_SEG), ("seg_len", C.c_int64 * MAX_SEG),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("grad_scale", C.c_float), ("step", C.c_int32),
                ("clamp_off", C.c_int64), ("clamp_len", C.c_int64), ("clamp_lo", C.c_float), ("clamp_hi", C.c_float),
                ("dev_step", C.c_void_p)]


class PendulumConfig(C.Structure):
    _fields_ = [("node", C.c_int32), ("n_dec", C.c_int32), ("factor", C.c_int32 * MAX_DEC),
                ("dec_extra", C.c_int32 * MAX_DEC), ("col_lo", C.c_int32 * MAX_DEC), ("col_hi", C.c_int32 * MAX_DEC), ("scm", C.c_int32),
                ("flow_num", C.c_int32), ("input_dim", C.c_int32), ("hidden", C.c_int32), ("gemm_mode", C.c_int32),
                ("general_mask", C.c_int32), ("n_params", C.c_int64), ("enc", Linear * 3), ("dec", (Linear * 3) * MAX_DEC),
                ("flow_off", C.c_int64 * MAX_NODE), ("I_B_inv", C.c_float * (MAX_NODE * MAX_NODE)),
                ("beta", C.c_float), ("lambda_", C.c_float)]


class PendulumIO(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("x", C.c_void_p), ("y", C.c_void_p), ("ld_y", C.c_int32),
                ("noise", C.c_void_p), ("batch", C.c_int64), ("x_l", C.c_void_p), ("y_l", C.c_void_p),
                ("ld_y_l", C.c_int32), ("batch_l", C.c_int64), ("logs", C.c_void_p), ("xhat", C.c_void_p),
                ("masks", C.c_void_p), ("d_params", C.c_void_p), ("d_grads", C.c_void_p), ("d_net", Linear * 3),
                ("d_n_params", C.c_int64), ("perm", C.c_void_p), ("gamma", C.c_float)]


class PendulumFwdIO(C.Structure):
    _fields_ = [("params", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("x", C.c_void_p), ("noise", C.c_void_p), ("latent_in", C.c_void_p), ("batch", C.c_int64),
                ("deterministic", C.c_int32), ("mean", C.c_void_p), ("logvar", C.c_void_p), ("epsilon", C.c_void_p),
                ("orig_latent", C.c_void_p), ("latent", C.c_void_p), ("align_latent", C.c_void_p),
                ("xhat_separated", C.c_void_p), ("xhat", C.c_void_p), ("masks", C.c_void_p)]


class TabularConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("node", C.c_int32), ("n_dec", C.c_int32), ("factor", C.c_int32 * MAX_DEC),
                ("out_dim", C.c_int32 * MAX_DEC), ("scm", C.c_int32), ("flow_num", C.c_int32),
                ("input_dim", C.c_int32), ("act", C.c_int32), ("n_enc_layers", C.c_int32),
                ("n_dec_layers", C.c_int32), ("enc", Linear * MAX_LAYERS), ("dec", (Linear * MAX_LAYERS) * MAX_DEC),
                ("flow_off", C.c_int64 * MAX_NODE), ("sigma_off", C.c_int64), ("flatten_topology", C.c_int32 * 16),
                ("n_span", C.c_int32), ("span_start", C.c_int32 * MAX_SPANS), ("span_dim", C.c_int32 * MAX_SPANS),
                ("span_kind", C.c_int32 * MAX_SPANS), ("n_params", C.c_int64),
                ("I_B_inv", C.c_float * (MAX_NODE * MAX_NODE)), ("beta", C.c_float), ("lambda_", C.c_float)]


class TabularIO(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("x", C.c_void_p), ("y", C.c_void_p), ("noise", C.c_void_p),
                ("batch", C.c_int64), ("logs", C.c_void_p), ("xhat", C.c_void_p), ("latents", C.c_void_p)]


N_GEN, N_GEN_BLOCKS, N_RES = 5, 5, 8


class Conv(C.Structure):
    _fields_ = [("w", C.c_int64), ("b", C.c_int64), ("u", C.c_int64), ("v", C.c_int64), ("cin", C.c_int32),
                ("cout", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32)]


class BNorm(C.Structure):
    _fields_ = [("weight", C.c_int64), ("bias", C.c_int64), ("running_mean", C.c_int64), ("running_var", C.c_int64),
                ("c", C.c_int32), ("momentum", C.c_float), ("eps", C.c_float)]


class GenBlock(C.Structure):
    _fields_ = [("conv1", Conv), ("conv2", Conv), ("conv0", Conv), ("bn1", BNorm), ("bn2", BNorm)]


class GeneratorDesc(C.Structure):
    _fields_ = [("z_dim", C.c_int32), ("z_src", C.c_int32 * MAX_NODE), ("lin0", Conv), ("blk", GenBlock * N_GEN_BLOCKS),
                ("attn", Conv * 4), ("bn", BNorm), ("to_rgb", Conv)]


class ResBlock(C.Structure):
    _fields_ = [("conv1", Conv), ("conv2", Conv), ("down", Conv), ("bn1", BNorm), ("bn2", BNorm), ("bn_down", BNorm),
                ("has_down", C.c_int32)]


class CelebaConfig(C.Structure):
    _fields_ = [("node", C.c_int32), ("latent_dim", C.c_int32), ("scm", C.c_int32), ("flow_num", C.c_int32),
                ("image_size", C.c_int32), ("gemm_mode", C.c_int32), ("n_params", C.c_int64), ("n_frozen", C.c_int64),
                ("fc", Linear), ("flow_off", C.c_int64 * MAX_NODE), ("I_B_inv", C.c_float * (MAX_NODE * MAX_NODE)),
                ("beta", C.c_float), ("lambda_", C.c_float), ("rn_conv1", Conv), ("rn_bn1", BNorm),
                ("rn_blk", ResBlock * N_RES), ("gen", GeneratorDesc * N_GEN)]


class CelebaIO(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("frozen", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("x", C.c_void_p), ("ld_x", C.c_int32), ("y", C.c_void_p),
                ("ld_y", C.c_int32), ("masks", C.c_void_p), ("noise1", C.c_void_p), ("noise2", C.c_void_p),
                ("batch", C.c_int64), ("backward", C.c_int32), ("deterministic", C.c_int32),
                ("encoder_passes", C.c_int32), ("logs", C.c_void_p), ("xhat", C.c_void_p),
                ("xhat_separated", C.c_void_p), ("latents", C.c_void_p), ("encode_only", C.c_int32),
                ("latent_in", C.c_void_p), ("epsilon2_in", C.c_void_p)]


MAX_TCOL, MAX_TCOMP, MAX_TCAT = 16, 10, 16


class TvaeColumn(C.Structure):
    _fields_ = [("kind", C.c_int32), ("out_start", C.c_int32), ("n_all", C.c_int32), ("n_valid", C.c_int32),
                ("valid_idx", C.c_int32 * MAX_TCOMP), ("round_int", C.c_int32), ("mean", C.c_double * MAX_TCOMP),
                ("std", C.c_double * MAX_TCOMP), ("prec", C.c_double * MAX_TCOMP), ("log_a", C.c_double * MAX_TCOMP),
                ("category", C.c_double * MAX_TCAT)]


class TvaeTransformConfig(C.Structure):
    _fields_ = [("n_col", C.c_int32), ("out_dim", C.c_int32), ("col", TvaeColumn * MAX_TCOL)]


# every symbol include/cdgvae.h declares
EXPORTS = ["cdg_last_error", "cdg_version", "cdg_device_ok", "cdg_launch_count", "cdg_launch_count_add", "cdg_abi_sizeof", "cdg_flow_apply", "cdg_pendulum_profile_enable",
           "cdg_pendulum_profile_read", "cdg_adam_step", "cdg_allreduce_oneshot", "cdg_pendulum_create",
           "cdg_pendulum_destroy", "cdg_pendulum_workspace_bytes", "cdg_pendulum_workspace_offset", "cdg_pendulum_workspace_bytes_infomax", "cdg_pendulum_forward_backward", "cdg_pendulum_ready_events_enable", "cdg_pendulum_ready_event", "cdg_stream_wait_event",
           "cdg_pendulum_forward", "cdg_tabular_create", "cdg_tabular_destroy", "cdg_tabular_workspace_bytes",
           "cdg_tabular_forward_backward", "cdg_tabular_forward", "cdg_tabular_const_params", "cdg_tabular_tvae_tile", "cdg_gemm", "cdg_celeba_create", "cdg_celeba_destroy",
           "cdg_celeba_workspace_bytes", "cdg_celeba_step", "cdg_celeba_generator_streams", "cdg_conv2d_workspace_bytes", "cdg_conv2d_forward",
           "cdg_conv2d_dgrad", "cdg_split_bf16", "cdg_gemm_bsplit", "cdg_gemm_planes", "cdg_gemm_planes_acc", "cdg_tvae_transform", "cdg_tvae_inverse_transform",
           "cdg_gumbel_argmax", "cdg_pixels_to_float", "cdg_pixels_gather_to_float"]

_lib = None
ABI_STRUCTS = [Linear, AdamArgs, PendulumConfig, PendulumIO, PendulumFwdIO, TabularConfig, TabularIO, Conv, BNorm, CelebaConfig,
               CelebaIO, TvaeTransformConfig]      # filled below, in the order cdg_abi_sizeof() enumerates them


def _check_abi(L, path):
    """Refuse a library whose struct layouts differ from the ctypes declarations in this file (a stale prebuilt .so)."""
    if not hasattr(L, "cdg_abi_sizeof"):
        raise RuntimeError(f"{path} predates this binding (no cdg_abi_sizeof): rebuild it")
    L.cdg_abi_sizeof.restype = C.c_int64
    L.cdg_abi_sizeof.argtypes = [C.c_int]
    for i, st in enumerate(ABI_STRUCTS):
        got = L.cdg_abi_sizeof(i)
        if got != C.sizeof(st):
            raise RuntimeError(f"{path}: sizeof({st.__name__}) is {got} in the library, {C.sizeof(st)} in this binding; "
                               "the library is stale -- rebuild it (python -m cdgvae_b200.build)")
    if L.cdg_abi_sizeof(len(ABI_STRUCTS)) != -1:
        raise RuntimeError(f"{path} declares more structs than this binding knows: the binding is stale")


def lib():
    """Load libcdgvae_sm100.so (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    exp = os.environ.get("CDG_EXPERIMENTS_LIB") == "1"     # development only: the build with the experiment switches live
    path = _build.LIB_EXP if exp else _build.LIB
    if _build.stale(exp):
        try:
            _build.build(experiments=exp)
        except _build.NoCompiler as e:  # no nvcc on the box: a prebuilt library must be present (its layout is checked below)
            if not os.path.exists(path):
                raise RuntimeError(f"libcdgvae_sm100.so is missing and cannot be built: {e}") from e
        # any other failure (compile / link error, lock or permission problem) propagates: an older library on disk may
        # not match the structures declared here
    if not os.path.exists(path):
        raise RuntimeError("libcdgvae_sm100.so is missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    _check_abi(L, path)
    L.cdg_last_error.restype = C.c_char_p
    L.cdg_launch_count.restype = C.c_longlong
    L.cdg_launch_count_add.argtypes = [C.c_longlong]
    L.cdg_launch_count_add.restype = None
    L.cdg_pendulum_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.cdg_pendulum_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.cdg_pendulum_workspace_bytes.restype = C.c_int64
    L.cdg_pendulum_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    L.cdg_pendulum_workspace_offset.restype = C.c_int64
    L.cdg_pendulum_workspace_offset.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
    L.cdg_pendulum_workspace_bytes_infomax.restype = C.c_int64
    L.cdg_pendulum_workspace_bytes_infomax.argtypes = [C.c_void_p, C.c_int64]
    L.cdg_pendulum_create.argtypes = [C.POINTER(PendulumConfig), C.POINTER(C.c_void_p)]
    L.cdg_pendulum_destroy.argtypes = [C.c_void_p]
    L.cdg_pendulum_destroy.restype = None
    L.cdg_pendulum_ready_events_enable.argtypes = [C.c_void_p, C.c_int]
    L.cdg_pendulum_ready_event.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.cdg_stream_wait_event.argtypes = [C.c_void_p, C.c_void_p]
    L.cdg_pendulum_forward_backward.argtypes = [C.c_void_p, C.POINTER(PendulumIO), C.c_void_p]
    L.cdg_pendulum_forward.argtypes = [C.c_void_p, C.POINTER(PendulumFwdIO), C.c_void_p]
    L.cdg_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(AdamArgs), C.c_void_p]
    L.cdg_allreduce_oneshot.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p]
    L.cdg_tabular_workspace_bytes.restype = C.c_int64
    L.cdg_tabular_workspace_bytes.argtypes = [C.c_void_p, C.c_int64]
    L.cdg_tabular_create.argtypes = [C.POINTER(TabularConfig), C.POINTER(C.c_void_p)]
    L.cdg_tabular_destroy.argtypes = [C.c_void_p]
    L.cdg_tabular_destroy.restype = None
    L.cdg_tabular_forward_backward.argtypes = [C.c_void_p, C.POINTER(TabularIO), C.c_void_p]
    L.cdg_tabular_forward.argtypes = [C.c_void_p, C.POINTER(TabularIO), C.c_int32, C.c_void_p]
    L.cdg_gemm.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                           C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    L.cdg_celeba_create.argtypes = [C.POINTER(CelebaConfig), C.POINTER(C.c_void_p)]
    L.cdg_celeba_destroy.argtypes = [C.c_void_p]
    L.cdg_celeba_destroy.restype = None
    L.cdg_celeba_workspace_bytes.restype = C.c_int64
    L.cdg_celeba_workspace_bytes.argtypes = [C.c_void_p, C.c_int64]
    L.cdg_celeba_step.argtypes = [C.c_void_p, C.POINTER(CelebaIO), C.c_void_p]
    L.cdg_conv2d_workspace_bytes.restype = C.c_int64
    L.cdg_conv2d_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(Conv), C.c_int32]
    L.cdg_conv2d_forward.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.POINTER(Conv), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.cdg_conv2d_dgrad.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(Conv),
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.cdg_split_bf16.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    L.cdg_gemm_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_void_p]
    L.cdg_pixels_gather_to_float.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.cdg_gemm_planes_acc.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.cdg_gemm_bsplit.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.cdg_tvae_transform.argtypes = [C.POINTER(TvaeTransformConfig), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int64, C.c_void_p]
    L.cdg_tvae_inverse_transform.argtypes = [C.POINTER(TvaeTransformConfig), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    L.cdg_gumbel_argmax.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.cdg_pixels_to_float.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.cdg_flow_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().cdg_last_error().decode()
        if rc == 1:
            raise ValueError(msg)
        raise RuntimeError(f"libcdgvae_sm100 error {rc}: {msg}")


def require_cuda(device):
    import torch
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cdgvae_b200 runs on sm_100a CUDA devices only (no CPU path); got device %r" % (device,))
    return dev
