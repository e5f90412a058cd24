"""ctypes binding of libsrdiff_b200.so (declared in include/srdiff_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails the
caller gets an exception.  The product path never runs on the CPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRD_LIB_PATH: developer override used to A/B kernel build variants on the GPU box
LIB_PATH = os.environ.get("SRD_LIB_PATH") or os.path.join(_HERE, "libsrdiff_b200.so")

SRD_MAX_SRC = 3
F32, BF16 = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TC = 0, 1, 2
ACT_NONE, ACT_MISH, ACT_LRELU02 = 0, 1, 2
BORDER_KEEP, BORDER_REFLECT = 0, 1
SRC_WIN8 = 1


class SrdSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ct", C.c_int32), ("c_off", C.c_int32),
                ("cin", C.c_int32), ("flags", C.c_int32)]


class SrdConvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("kernel", C.c_int32), ("n_src", C.c_int32),
        ("B", C.c_int32), ("H_in", C.c_int32), ("W_in", C.c_int32),
        ("src", SrdSrc * SRD_MAX_SRC),
        ("stride", C.c_int32), ("nky", C.c_int32), ("nkx", C.c_int32),
        ("ky0", C.c_int32), ("kx0", C.c_int32), ("up", C.c_int32), ("N", C.c_int32),
        ("w_direct", C.c_void_p), ("w_tc", C.c_void_p), ("bias", C.c_void_p),
        ("act", C.c_int32), ("alpha", C.c_float),
        ("temb", C.c_void_p), ("temb_bstride", C.c_int32),
        ("res1", C.c_void_p), ("res1_ct", C.c_int32), ("res1_coff", C.c_int32), ("beta1", C.c_float),
        ("res2", C.c_void_p), ("res2_ct", C.c_int32), ("res2_coff", C.c_int32), ("beta2", C.c_float),
        ("out", C.c_void_p), ("out_ct", C.c_int32), ("out_coff", C.c_int32),
        ("out2", C.c_void_p), ("out2_ct", C.c_int32), ("out2_coff", C.c_int32),
        ("H_out", C.c_int32), ("W_out", C.c_int32), ("border", C.c_int32),
        ("planar", C.c_void_p), ("planar_c", C.c_int32), ("planar_clamp01", C.c_int32),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p),
        ("lin_src", C.c_void_p), ("lin_ct", C.c_int32), ("lin_c", C.c_int32), ("lin_w", C.c_void_p),
        ("N2", C.c_int32), ("dst2_ct", C.c_int32), ("dst2_coff", C.c_int32), ("reserved2", C.c_int32),
        ("bias2", C.c_void_p), ("dst2", C.c_void_p),
        ("reverse_items", C.c_int32), ("n_split", C.c_int32),
    ]


class SrdPSampleArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("eps", C.c_void_p), ("noise", C.c_void_p),
        ("coef", C.c_void_p), ("t", C.c_void_p),
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("clip_denoised", C.c_int32),
        ("seed", C.c_uint64), ("step", C.c_uint32), ("image_offset", C.c_uint32),
        ("x_out", C.c_void_p), ("x0_out", C.c_void_p),
        ("img_out", C.c_void_p), ("img_lr_up", C.c_void_p), ("res_rescale", C.c_float),
        ("clip_input", C.c_int32),
        ("nhwc_out", C.c_void_p), ("nhwc_ct", C.c_int32), ("act_dtype", C.c_int32),
        ("seed_ptr", C.c_void_p),
    ]


class SrdChainLink(C.Structure):
    _fields_ = [("src_layer", C.c_int32 * SRD_MAX_SRC), ("src_which", C.c_int32 * SRD_MAX_SRC),
                ("res1_layer", C.c_int32), ("res1_which", C.c_int32), ("temb_off", C.c_int32)]


class SrdGroupNormArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_ct", C.c_int32), ("src_coff", C.c_int32),
        ("dtype", C.c_int32),
        ("out", C.c_void_p), ("out_ct", C.c_int32), ("out_coff", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("groups", C.c_int32),
        ("act", C.c_int32), ("border", C.c_int32),
        ("eps", C.c_float), ("temb_bstride", C.c_int32),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("temb", C.c_void_p),
        ("res", C.c_void_p), ("res_ct", C.c_int32), ("res_coff", C.c_int32),
        ("reserved", C.c_int32),
        ("stats", C.c_void_p),
    ]


# every symbol include/srdiff_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f, _u32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_uint64
SYMBOLS = {
    "srd_last_error": (C.c_char_p, []),
    "srd_version": (C.c_int, []),
    "srd_abi_sizeof": (C.c_int, [_i32]),
    "srd_conv_tc_supported": (C.c_int, [C.POINTER(SrdConvDesc)]),
    "srd_conv_plan_create": (C.c_int, [C.POINTER(SrdConvDesc), C.POINTER(_vp)]),
    "srd_conv_plan_destroy": (C.c_int, [_vp]),
    "srd_conv_plan_kernel": (C.c_int, [_vp]),
    "srd_conv_run": (C.c_int, [_vp, _vp, _vp]),
    "srd_watchdog_flag": (C.c_int, []),
    "srd_chain_create": (C.c_int, [C.POINTER(_vp), C.POINTER(SrdChainLink), _i32, _i32, C.POINTER(_vp)]),
    "srd_chain_run": (C.c_int, [_vp, _vp, _vp]),
    "srd_chain_destroy": (C.c_int, [_vp]),
    "srd_chain_segments": (C.c_int, [_vp]),
    "srd_chain_max_layers": (C.c_int, []),
    "srd_chain_fused_max_layers": (C.c_int, []),
    "srd_conv_plans_share_segment": (C.c_int, [_vp, _vp]),
    "srd_tc_set_mode": (C.c_int, [_i32, _i32]),
    "srd_tc_s2_halo": (C.c_int, []),
    "srd_tc_set_debug": (C.c_int, [_vp]),
    "srd_time_table": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "srd_sinusoidal_emb": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "srd_p_sample_step": (C.c_int, [C.POINTER(SrdPSampleArgs), _vp]),
    "srd_randn": (C.c_int, [_vp, _i32, _i32, _u64, _u32, _u32, _vp]),
    "srd_nchw_to_act": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _f, _f, _vp, _i32, _i32, _i32, _vp]),
    "srd_act_to_nchw": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "srd_upsample_nearest2x": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "srd_mish_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "srd_res2img": (C.c_int, [_vp, _vp, _vp, _i64, _f, _i32, _i32, _vp]),
    "srd_q_sample": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "srd_axpby": (C.c_int, [_vp, _vp, _f, _f, _i64, _vp, _vp]),
    "srd_bicubic_upsample": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "srd_u8hwc_to_nchw": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "srd_nchw_to_u8hwc": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "srd_tiles": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "srd_groupnorm_act": (C.c_int, [C.POINTER(SrdGroupNormArgs), _vp]),
    "srd_linear_attention": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp]),
    "srd_linear_f32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "srd_ssim_workspace": (C.c_int, [_i32, _i32, _i32, _i32]),
    "srd_ssim": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _f, _f, _vp, _vp, _vp]),
    "srd_resize_dim": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "srd_ssim_uniform_u8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "srd_f32_to_u8_round": (C.c_int, [_vp, _i64, _vp, _vp]),
    "srd_sq_err_u8": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
}

_lib = None


class SrdError(RuntimeError):
    pass


def lib():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SrdError(
                "libsrdiff_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C dgm_img_super_resolution_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        abi = (C.sizeof(SrdConvDesc), C.sizeof(SrdPSampleArgs), SrdConvDesc.dst2.offset,
               SrdPSampleArgs.seed_ptr.offset, C.sizeof(SrdGroupNormArgs), SrdGroupNormArgs.stats.offset)
        got = tuple(L.srd_abi_sizeof(i) for i in range(6))
        if abi != got:
            raise SrdError("ctypes structs disagree with include/srdiff_b200.h: %s vs %s" % (abi, got))
        # tuning / validation switches of the tcgen05 kernel (see srd_tc_set_mode in the header)
        if "SRD_TC_A_MODE" in os.environ or "SRD_TC_B_RESIDENT" in os.environ:
            if L.srd_tc_set_mode(int(os.environ.get("SRD_TC_A_MODE", "1")), int(os.environ.get("SRD_TC_B_RESIDENT", "1"))):
                raise SrdError("srd_tc_set_mode rejected SRD_TC_A_MODE / SRD_TC_B_RESIDENT: %s" % L.srd_last_error().decode())
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().srd_last_error()
        wd = lib().srd_watchdog_flag()
        raise SrdError("%s failed: %s%s" % (what or "libsrdiff_b200 call", msg.decode() if msg else "?",
                                            " [tcgen05 pipeline watchdog fired: wait #%d]" % wd if wd else ""))
