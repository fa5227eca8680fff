"""ctypes binding of libgddim_b200.so (include/gddim_b200.h).

The library is the product: there is no Python/CPU fallback for the hot path.  If the shared object is
missing this module raises at import of the symbol table; if no CUDA device is present, every compute
entry point fails with RuntimeError (host-side table functions still work).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgddim_b200.so")


class ModelCfg(C.Structure):
  _fields_ = [("image_size", C.c_int), ("data_channels", C.c_int), ("state_mult", C.c_int), ("nf", C.c_int),
              ("n_levels", C.c_int), ("ch_mult", C.c_int * 8), ("num_res_blocks", C.c_int), ("n_attn", C.c_int),
              ("attn_resolutions", C.c_int * 8), ("fir", C.c_int), ("skip_rescale", C.c_int),
              ("progressive_input", C.c_int), ("embedding_type", C.c_int), ("conditional", C.c_int),
              ("centered", C.c_int)]


class SamplerCfg(C.Structure):
  _fields_ = [("kind", C.c_int), ("nfe", C.c_int), ("deis_order", C.c_int), ("ts_order", C.c_int),
              ("denoising", C.c_int), ("mixed_score", C.c_int), ("use_graph", C.c_int),
              ("x_mul", C.c_float), ("x_add", C.c_float), ("lambda_coef", C.c_float), ("sdeis_use_order0", C.c_int),
              ("seed", C.c_ulonglong)]


class GemmDesc(C.Structure):
  _fields_ = [("a0", C.c_void_p), ("a0_ctot", C.c_int), ("a0_coff", C.c_int), ("a0_c", C.c_int), ("a0_taps", C.c_int),
              ("a1", C.c_void_p), ("a1_ctot", C.c_int), ("a1_coff", C.c_int), ("a1_c", C.c_int), ("a1_taps", C.c_int),
              ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
              ("w", C.c_void_p), ("N", C.c_int), ("w_ld", C.c_int), ("w_koff", C.c_int),
              ("w_batch_stride", C.c_longlong), ("w_rows_per_batch", C.c_int),
              ("bias", C.c_void_p), ("bias2", C.c_void_p), ("residual", C.c_void_p), ("rowscale", C.c_void_p),
              ("scale", C.c_float), ("out32", C.c_void_p), ("out16", C.c_void_p), ("row_out", C.c_void_p),
              ("ldo", C.c_int), ("epi", C.c_int), ("impl", C.c_int), ("force_block_n", C.c_int),
              ("force_m_sub", C.c_int), ("n_store", C.c_int), ("force_cta_pairs", C.c_int), ("reverse", C.c_int),
              ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p), ("gn_eps", C.c_float), ("gn_groups", C.c_int),
              ("gn_silu", C.c_int), ("wsplit", C.c_int)]


class NormDesc(C.Structure):
  _fields_ = [("src1", C.c_void_p), ("c1", C.c_int), ("src2", C.c_void_p), ("c2", C.c_int),
              ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("groups", C.c_int),
              ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
              ("silu", C.c_int), ("resample", C.c_int), ("dst16", C.c_void_p), ("raw16", C.c_void_p),
              ("raw_scale", C.c_float), ("reverse", C.c_int)]


class Step(C.Structure):
  _fields_ = [("t", C.c_double), ("n_eps", C.c_int), ("first_eps", C.c_int), ("A", C.c_float * 4),
              ("C", (C.c_float * 4) * 6), ("F", C.c_float * 4), ("M", C.c_float * 4), ("trace", C.c_int),
              ("has_P", C.c_int), ("P", C.c_float * 4)]


CTX_PRECISE_WEIGHTS = 1
CLD_DEIS, CLD_ORDER0, BLUR_ORDER0, CLD_SDEIS, CLD_PROGRAM = 0, 1, 2, 3, 4

_P = C.c_void_p
_D = C.POINTER(C.c_double)

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header is exported
SIGNATURES = {
    "gddim_last_error": (C.c_char_p, []),
    "gddim_abi_version": (C.c_int, []),
    "gddim_cuda_available": (C.c_int, []),
    "gddim_ctx_create": (C.c_int, [C.c_int, C.POINTER(ModelCfg), C.c_int, C.POINTER(_P)]),
    "gddim_ctx_create_ex": (C.c_int, [C.c_int, C.POINTER(ModelCfg), C.c_int, C.c_uint, C.POINTER(_P)]),
    "gddim_ctx_destroy": (None, [_P]),
    "gddim_param_count": (C.c_int, [_P]),
    "gddim_param_spec": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int * 4), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "gddim_param_set": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t]),
    "gddim_ctx_finalize": (C.c_int, [_P]),
    "gddim_ctx_set_gemm_impl": (C.c_int, [_P, C.c_int]),
    "gddim_ctx_workspace_bytes": (C.c_size_t, [_P]),
    "gddim_ctx_launch_count": (C.c_longlong, [_P]),
    "gddim_ctx_plan_size": (C.c_int, [_P]),
    "gddim_ctx_plan_op": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "gddim_ctx_set_profile": (C.c_int, [_P, C.c_int]),
    "gddim_ctx_get_profile": (C.c_int, [_P, _P, _P, _P]),
    "gddim_ctx_dump_profile": (C.c_int, [_P, C.c_char_p]),
    "gddim_ctx_get_profile_hbm": (C.c_int, [_P, _P]),
    "gddim_unet_forward": (C.c_int, [_P, _P, C.c_float, _P, C.c_int, _P]),
    "gddim_cld_create": (C.c_int, [C.c_double] * 6 + [C.c_int, C.POINTER(_P)]),
    "gddim_cld_destroy": (None, [_P]),
    "gddim_cld_R": (C.c_int, [_P, _P, C.c_int, _P]),
    "gddim_cld_psi": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "gddim_cld_F": (C.c_int, [_P, C.c_double, _P]),
    "gddim_cld_G": (C.c_int, [_P, C.c_double, _P]),
    "gddim_cld_eps_integrand": (C.c_int, [_P, _P, C.c_int, _P]),
    "gddim_cld_deis_coef": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "gddim_cld_order0_coef": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "gddim_cld_sdeis_coef": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, _P, C.c_int, _P]),
    "gddim_mvn_factor_svd": (C.c_int, [_P, _P]),
    "gddim_rev_ts": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_int, _P]),
    "gddim_blur_create": (C.c_int, [C.c_double, C.c_double, C.POINTER(_P)]),
    "gddim_blur_destroy": (None, [_P]),
    "gddim_blur_sampling_T": (C.c_double, [_P]),
    "gddim_blur_y_mean_coef": (C.c_int, [_P, C.c_double, _P]),
    "gddim_blur_y_std_coef": (C.c_double, [_P, C.c_double]),
    "gddim_blur_t2alpha": (C.c_double, [_P, C.c_double]),
    "gddim_multistep_ab_step": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_longlong, _P]),
    "gddim_scalar_ab_step": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_longlong, _P]),
    "gddim_relayout": (C.c_int, [_P, _P, C.c_longlong, C.c_int, C.c_int, _P]),
    "gddim_dct2d_32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "gddim_conv_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "gddim_gemm_gnf_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gddim_group_norm": (C.c_int, [C.POINTER(NormDesc), _P]),
    "gddim_attention": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "gddim_gn_qkv": (C.c_int, [_P, _P, _P, C.c_int, C.c_float, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "gddim_attention_proj": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, _P]),
    "gddim_sampler_create": (C.c_int, [_P, C.POINTER(SamplerCfg), _P, _P, C.POINTER(_P)]),
    "gddim_sampler_create_ts": (C.c_int, [_P, C.POINTER(SamplerCfg), _P, _P, _P, C.c_int, C.POINTER(_P)]),
    "gddim_sampler_create_program": (C.c_int, [_P, C.POINTER(SamplerCfg), _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "gddim_cld_ldeis_coef": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "gddim_cld_mldeis_coef": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "gddim_cld_psi1": (C.c_int, [_P, C.c_double, C.c_int, _P]),
    "gddim_sampler_destroy": (None, [_P]),
    "gddim_sampler_alive": (C.c_int, [_P]),
    "gddim_sampler_set_seed": (C.c_int, [_P, C.c_ulonglong]),
    "gddim_sampler_coef": (C.c_longlong, [_P, _P, C.c_longlong]),
    "gddim_sampler_num_steps": (C.c_int, [_P]),
    "gddim_sampler_rev_ts": (C.c_int, [_P, _P, C.c_int]),
    "gddim_sample": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "gddim_sample_noise": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "gddim_sampler_launch_count": (C.c_longlong, [_P]),
    "gddim_sampler_time_update": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
}

_lib = None


def lib():
  """Loads the shared library (once).  Raises if it has not been built: run `make` or __graft_entry__.build()."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(f"{LIB_PATH} not found: build it with `make` (nvcc, sm_100a). "
                         "gddim_b200 has no CPU fallback for the sampling hot path.")
    h = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(h, name)            # AttributeError if the library does not export a declared symbol
      fn.restype, fn.argtypes = res, args
    _lib = h
  return _lib


def last_error():
  return lib().gddim_last_error().decode()


def check(rc, what=""):
  if rc != 0:
    raise RuntimeError(f"{what}: {last_error()}" if what else last_error())


def cuda_available():
  return bool(lib().gddim_cuda_available())


def require_cuda(what):
  if not cuda_available():
    raise RuntimeError(f"{what} needs a CUDA device (sm_100a); gddim_b200 has no CPU fallback")
