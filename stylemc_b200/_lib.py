"""ctypes binding of libstylemc_b200.so (the C ABI declared in include/stylemc_b200.h).

This is the FFI boundary that replaces the reference's ``custom_ops.get_plugin`` + pybind modules
(torch_utils/custom_ops.py:46-124; upfirdn2d.py:26-35; bias_act.py:41-52).  Differences on purpose:
the library is built ahead of time (stylemc_b200/build.py), and a missing library or a non-zero status
raises -- there is no fallback to a slow reference implementation (upfirdn2d.py:33-35 warns and falls
back; we never do).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstylemc_b200.so')

F32, F16, F64 = 0, 1, 2
DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.float64: F64}
MAX_TAPS = 32

_STATUS = {-1: 'invalid argument', -2: 'unsupported configuration', -3: 'tensor too large', -4: 'CUDA driver entry point / tensor map failure'}


class Tap(ctypes.Structure):
    _fields_ = [('dn', ctypes.c_int32), ('dy', ctypes.c_int32), ('dx', ctypes.c_int32), ('brow', ctypes.c_int32)]


class Epilogue(ctypes.Structure):
    _fields_ = [('row_scale', ctypes.c_void_p), ('post_scale', ctypes.c_void_p), ('bias', ctypes.c_void_p),
                ('noise', ctypes.c_void_p), ('noise_sh', ctypes.c_int64), ('noise_sw', ctypes.c_int64),
                ('act', ctypes.c_int32), ('alpha', ctypes.c_float), ('gain', ctypes.c_float), ('clamp', ctypes.c_float),
                ('residual', ctypes.c_void_p), ('out_f32', ctypes.c_void_p), ('out_hi', ctypes.c_void_p),
                ('out_lo', ctypes.c_void_p), ('out_raw', ctypes.c_void_p),
                ('o_sn', ctypes.c_int64), ('o_sh', ctypes.c_int64), ('o_sw', ctypes.c_int64), ('o_off', ctypes.c_int64),
                ('acc_scale', ctypes.c_float),
                ('out_raw_lo', ctypes.c_void_p), ('rgb_w', ctypes.c_void_p), ('rgb_acc', ctypes.c_void_p),
                ('rgb_sn', ctypes.c_int64), ('rgb_sj', ctypes.c_int64), ('rgb_sh', ctypes.c_int64),
                ('mask_y', ctypes.c_void_p), ('mask_y_lo', ctypes.c_void_p), ('mask_grgb', ctypes.c_void_p), ('rgb_snt', ctypes.c_int64)]


class IgemmDesc(ctypes.Structure):
    _fields_ = [('A', ctypes.c_void_p), ('NA', ctypes.c_int32), ('HA', ctypes.c_int32), ('WA', ctypes.c_int32),
                ('C', ctypes.c_int32), ('lda', ctypes.c_int64),
                ('B', ctypes.c_void_p), ('rowsB', ctypes.c_int32), ('ldb', ctypes.c_int64),
                ('n_img', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('n_out', ctypes.c_int32),
                ('tw', ctypes.c_int32), ('th', ctypes.c_int32), ('tn', ctypes.c_int32), ('ntaps', ctypes.c_int32),
                ('taps', Tap * MAX_TAPS), ('epi', Epilogue), ('acc_chunk_k', ctypes.c_int32),
                ('nprob', ctypes.c_int32), ('prob_ntaps', ctypes.c_int32 * 4), ('prob_o_off', ctypes.c_int64 * 4)]


class IgemmPlanInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('kernel', 'bn', 'kc', 'mode', 'Wt', 'Wp', 'RB', 'na_hi', 'na_lo', 'nb', 'b_resident', 'a_share', 'nprob',
                                              'kchunks', 'super_tiles', 'grid', 'smem_bytes', 'tmem_cols')] + \
               [(n, ctypes.c_int32 * 4) for n in ('prob_ntaps', 'prob_nsegs', 'prob_ndrains', 'prob_commits', 'prob_stages')] + [('pair', ctypes.c_int32)]


class UpfirdnParams(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int32), ('C', ctypes.c_int32), ('inH', ctypes.c_int32), ('inW', ctypes.c_int32),
                ('outH', ctypes.c_int32), ('outW', ctypes.c_int32),
                ('x_stride', ctypes.c_int64 * 4), ('y_stride', ctypes.c_int64 * 4),
                ('fH', ctypes.c_int32), ('fW', ctypes.c_int32), ('f_stride', ctypes.c_int64 * 2),
                ('upx', ctypes.c_int32), ('upy', ctypes.c_int32), ('downx', ctypes.c_int32), ('downy', ctypes.c_int32),
                ('padx0', ctypes.c_int32), ('pady0', ctypes.c_int32), ('flip', ctypes.c_int32), ('gain', ctypes.c_float),
                ('separable', ctypes.c_int32), ('fsep', ctypes.c_float * 8)]


_T = {'p': ctypes.c_void_p, 'i': ctypes.c_int, 'q': ctypes.c_int64, 'f': ctypes.c_float}

# name -> argument kinds, in header order (p pointer, i int32, q int64, f float)
SIGNATURES = {
    'smc_abi_version': '',
    'smc_bias_act': 'pppppp iqiq ii fff p',
    'smc_upfirdn2d': 'ppp i p p',
    'smc_igemm': 'pp',
    'smc_igemm_plan': 'pp',
    'smc_igemm_config': 'ii',
    'smc_synth_config': 'ii',
    'smc_demod_coefs': 'pp q p iii p',
    'smc_pack_nhwc': 'p q p q pp iiii p',
    'smc_unpack_nchw': 'p i pp iiii p',
    'smc_fir_act': 'p i iiii pppp fff p q pppp p',
    'smc_img_finish': 'ppp f p iii p i q p',
    'smc_torgb': 'pp iiii pp q f p f ppp p q pp p',
    'smc_act_bwd': 'pp iiii p i p q ppp q f p f pppp fff pppp p',
    'smc_fir_bwd': 'pp iiii pppp p',
    'smc_sgrad_finish': 'ppppp q pp iii p q p',
    'smc_grad_scale': 'p q f pp p',
    'smc_mask_scale': 'pppp q p',
    'smc_resample_fwd': 'pppppp i iii i pp p',
    'smc_resample_bwd': 'ppppppp i iii i p p p',
    'smc_patchify': 'ppp iii p',
    'smc_unpatchify': 'pp iii p',
    'smc_assemble_tokens': 'pppp iii p',
    'smc_embed_text': 'pppp iii p',
    'smc_layernorm_fwd': 'p qq pp pppp p q i p',
    'smc_layernorm_bwd': 'pp qq ppp p q ii p',
    'smc_attention_fwd': 'pppp iiiii p',
    'smc_attention_bwd': 'pppp iiiii p',
    'smc_attention_bwd_tiled': 'ppppp iiiii p',
    'smc_quickgelu_fwd': 'ppp q p',
    'smc_quickgelu_bwd': 'pppp q p',
    'smc_split_rows': 'ppp q iiii p',
    'smc_head_proj': 'ppp iii p',
    'smc_head_proj_bwd': 'ppp iii p',
    'smc_clip_loss': 'ppppp ii ff p f ii p',
    'smc_img_to_uint8': 'pp iiiii p',
    'smc_prepare_weights': 'p iiiii p pppp p p',
    'smc_prelu': 'pppp q ii p',
    'smc_adaptive_avg_pool': 'pp q iiiiiiii i p',
    'smc_pixelnorm': 'ppp iii p',
    'smc_adam_step': 'pppp q ffffff p',
    'smc_matmul_nt_f64acc': 'p qq p qq pp iii p',
    'smc_fma': 'pppp i pppp p',
    'smc_fma_reduce': 'ppp i pppp p',
    'smc_sgd_step': 'pp q fff p',
    'smc_sgd_step_dev': 'pp q p ff p',
}

_lib = None


def lib():
    """Load the shared library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -m stylemc_b200.build` (there is no CPU/torch fallback)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, sig in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_T[c] for c in sig.replace(' ', '')]
        if handle.smc_abi_version() != 1:
            raise RuntimeError('libstylemc_b200.so ABI version mismatch; rebuild')
        _lib = handle
        for key, env in ((0, 'STYLEMC_HCONV'), (2, 'STYLEMC_HCONV_NB'), (3, 'STYLEMC_HCONV_WT'), (4, 'STYLEMC_HCONV_GRID'), (5, 'STYLEMC_HCONV_MASK'), (6, 'STYLEMC_HCONV_MINPOS'),
                         (7, 'STYLEMC_HCONV_PAIR'), (8, 'STYLEMC_HCONV_EPI')):
            if os.environ.get(env):          # diagnostics only: A/B the halo-tile conv kernel against the per-tap kernel
                handle.smc_igemm_config(key, int(os.environ[env]))
        for key, env in ((0, 'STYLEMC_FIR_ACT3'), (1, 'STYLEMC_FIR_BWD3'), (2, 'STYLEMC_ACT_BWD2'), (3, 'STYLEMC_UPFIRDN_ROWS'), (4, 'STYLEMC_RESAMPLE_VFIRST'),
                         (5, 'STYLEMC_ATTENTION_TILED'), (6, 'STYLEMC_RESAMPLE_ROWS')):
            if os.environ.get(env):          # diagnostics only: A/B the newer glue kernels against the older ones
                handle.smc_synth_config(key, int(os.environ[env]))
    return _lib


def check(status, what):
    if status != 0:
        msg = _STATUS.get(status)
        if msg is None and status > 0:
            msg = f'CUDA error {status}'
        raise RuntimeError(f'{what}: {msg} (status {status})')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must reside on a CUDA device: stylemc_b200 has no CPU path')


# kernels launched per entry point (for bench.py's gpu_launches claim); everything else launches one
_LAUNCHES = {'smc_abi_version': 0, 'smc_igemm_config': 0, 'smc_igemm_plan': 0, 'smc_synth_config': 0, 'smc_resample_fwd': 2, 'smc_resample_bwd': 2, 'smc_grad_scale': 2,
             'smc_attention_bwd_tiled': 2}
launch_count = 0
igemm_hook = None      # bench.py installs a callable(desc_addr) -> context manager to time every smc_igemm launch


def call(name, *args):
    global launch_count
    launch_count += _LAUNCHES.get(name, 1)
    check(getattr(lib(), name)(*args), name)
