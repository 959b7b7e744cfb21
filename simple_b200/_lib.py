"""ctypes binding of the C ABI declared in include/simple_b200.h (built in-tree as csrc/libsimple_b200.so).

No CPU fallback: `lib()` raises if the shared library is missing or cannot be loaded.
"""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
_SO = os.path.join(_CSRC, 'libsimple_b200.so')
_LIB = None

SB_MAX_TAPS = 16


NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--use_fast_math',
              '-Xcompiler', '-fPIC']


def build(verbose=False, force=False):
    """Compile every CUDA source for sm_100a into csrc/libsimple_b200.so (nvcc cross-compiles without a GPU).
    One object file per source (compiled in parallel, only when stale), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = sorted(glob.glob(os.path.join(_CSRC, '*.cu')))
    common = glob.glob(os.path.join(_CSRC, '*.cuh')) + [os.path.join(_HERE, '..', 'include', 'simple_b200.h')]
    newest_common = max(os.path.getmtime(d) for d in common)
    objdir = os.path.join(_CSRC, 'build')
    os.makedirs(objdir, exist_ok=True)
    jobs, objs = [], []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_common):
            cmd = ['nvcc'] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for r in ex.map(lambda c: subprocess.run(c, cwd=_CSRC, capture_output=not verbose, text=True), jobs):
                if r.returncode != 0:
                    raise RuntimeError('nvcc failed:\n' + (r.stderr or '') + (r.stdout or ''))
    if jobs or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(o) for o in objs):
        subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', _SO] + objs, check=True,
                       cwd=_CSRC)
    return _SO


class TapGemmArgs(ctypes.Structure):
    _fields_ = [
        ('a', ctypes.c_void_p), ('B', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int), ('C', ctypes.c_int),
        ('Ho', ctypes.c_int), ('Wo', ctypes.c_int),
        ('bw', ctypes.c_int), ('bh', ctypes.c_int), ('bb', ctypes.c_int),
        ('ntaps', ctypes.c_int), ('tap_dy', ctypes.c_int * SB_MAX_TAPS), ('tap_dx', ctypes.c_int * SB_MAX_TAPS),
        ('w', ctypes.c_void_p), ('N', ctypes.c_int), ('ldw', ctypes.c_int),
        ('out', ctypes.c_void_p), ('out_f32', ctypes.c_int), ('ldo', ctypes.c_int),
        ('bias', ctypes.c_void_p), ('relu', ctypes.c_int), ('splits', ctypes.c_int), ('accumulate', ctypes.c_int),
        ('bias_mod', ctypes.c_int),
    ]


class TapWgradArgs(ctypes.Structure):
    _fields_ = [
        ('a', ctypes.c_void_p), ('B', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int), ('C', ctypes.c_int),
        ('dy', ctypes.c_void_p), ('Ho', ctypes.c_int), ('Wo', ctypes.c_int), ('N', ctypes.c_int),
        ('bw', ctypes.c_int), ('bh', ctypes.c_int), ('bb', ctypes.c_int),
        ('ntaps', ctypes.c_int), ('tap_dy', ctypes.c_int * SB_MAX_TAPS), ('tap_dx', ctypes.c_int * SB_MAX_TAPS),
        ('dw', ctypes.c_void_p), ('ldw', ctypes.c_int), ('splits', ctypes.c_int), ('dbias', ctypes.c_void_p),
    ]


class InjGrads(ctypes.Structure):
    _fields_ = [('dsc', ctypes.c_void_p * 9), ('dsh', ctypes.c_void_p * 9), ('sc_sb', ctypes.c_int * 9),
                ('sc_sc', ctypes.c_int * 9), ('sh_sb', ctypes.c_int * 9), ('sh_sc', ctypes.c_int * 9)]


class SbView(ctypes.Structure):
    _fields_ = [('p', ctypes.c_void_p), ('kind', ctypes.c_int), ('s', ctypes.c_int), ('pad', ctypes.c_int),
                ('H2', ctypes.c_int), ('W2', ctypes.c_int)]


class PrepArgs(ctypes.Structure):
    _fields_ = [
        ('B', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int), ('C', ctypes.c_int),
        ('inp', SbView), ('add', ctypes.c_void_p), ('relu_in', ctypes.c_int), ('in_f32', ctypes.c_int), ('add_f32', ctypes.c_int),
        ('sums', ctypes.c_void_p), ('eps', ctypes.c_float), ('gamma', ctypes.c_void_p), ('beta', ctypes.c_void_p),
        ('Ta', ctypes.c_void_p), ('out1', ctypes.c_void_p),
        ('p', ctypes.c_float), ('seed', ctypes.c_ulonglong), ('seed_ptr', ctypes.c_void_p), ('stream_id', ctypes.c_uint),
        ('keep', ctypes.c_void_p), ('Tb', ctypes.c_void_p), ('sc', ctypes.c_void_p), ('sh', ctypes.c_void_p),
        ('out2', SbView), ('g2', SbView), ('g1', ctypes.c_void_p), ('bsums', ctypes.c_void_p),
        ('d_main', SbView), ('d_add', ctypes.c_void_p), ('dcol', ctypes.c_void_p),
        ('b_base', ctypes.c_int), ('b_count', ctypes.c_int), ('dgb', ctypes.c_void_p), ('moments', ctypes.c_int),
    ]


def lib():
    """Load the CUDA library; raise loudly if it is not there (the product has no fallback path)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise RuntimeError(
            f'simple_b200: CUDA extension {_SO} is missing. Build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` (needs nvcc). There is no CPU or PyTorch fallback.')
    L = ctypes.CDLL(_SO)
    L.sb_init.argtypes = [ctypes.c_int]
    L.sb_init.restype = ctypes.c_int
    L.sb_launch_count.restype = ctypes.c_longlong
    L.sb_version.restype = ctypes.c_char_p
    L.sb_tap_wgrad_bias_ok.argtypes = [ctypes.c_int, ctypes.c_int]
    L.sb_tap_wgrad_bias_ok.restype = ctypes.c_int
    L.sb_set_pair_mode.argtypes = [ctypes.c_int]
    L.sb_set_pair_mode.restype = ctypes.c_int
    L.sb_set_wgrad_mt2.argtypes = [ctypes.c_int]
    L.sb_set_wgrad_mt2.restype = ctypes.c_int
    L.sb_set_sm_reserve.argtypes = [ctypes.c_int]
    L.sb_dense_splits.argtypes = [ctypes.c_int] * 3
    L.sb_dense_splits.restype = ctypes.c_int
    L.sb_set_sm_reserve.restype = ctypes.c_int
    L.sb_set_wgrad_tap_groups.argtypes = [ctypes.c_int]
    L.sb_set_wgrad_tap_groups.restype = ctypes.c_int
    L.sb_set_ce_tma.argtypes = [ctypes.c_int]
    L.sb_set_ce_tma.restype = ctypes.c_int
    L.sb_set_prep_specialize.argtypes = [ctypes.c_int]
    L.sb_set_prep_specialize.restype = ctypes.c_int
    for name in dir(_Sigs):
        if name.startswith('sb_'):
            fn = getattr(L, name)
            fn.argtypes = getattr(_Sigs, name)
            fn.restype = ctypes.c_int
    _LIB = L
    return L


_p, _i, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong


class _Sigs:
    sb_tap_gemm = [ctypes.POINTER(TapGemmArgs), _p]
    sb_tap_wgrad = [ctypes.POINTER(TapWgradArgs), _p]
    sb_pixel_ce = [_p, _p, _i, _i, _f, _f, _p, _p, _p, _p, _p]
    sb_logits_ce = [ctypes.POINTER(TapGemmArgs), _p, _f, _f, _p, _p, _p]
    sb_stats = [ctypes.POINTER(PrepArgs), _p, _p]
    sb_prep = [ctypes.POINTER(PrepArgs), _p]
    sb_prep_bwd_reduce = [ctypes.POINTER(PrepArgs), _p]
    sb_prep_bwd_apply = [ctypes.POINTER(PrepArgs), _p]
    sb_standardize = [_p, _i, _i, _i, _p, _i, _i, _p, _p, _p, _i, _i, _p]
    sb_gru_blend = [_p, _p, _p, _p, _i, _ll, _p]
    sb_frame_feedback = [_p, _p, _p, _p, _f, _p, ctypes.c_ulonglong, _i, _i, _i, _i, _p]
    sb_gru_blend_bwd = [_p, _p, _i, _p, _i, _p, _ll, _p]
    sb_pack_conv = [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]
    sb_unpack_conv_grad = [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]
    sb_mean_attention_fwd = [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]
    sb_mean_attention_bwd = [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p]
    sb_lstm_sample2 = [_p] * 9 + [_i, _i, _p, _p, _p]
    sb_lstm_teacher_fwd = [_p, _p, _p, _p, _i, _i, _p, _p, _p, _p]
    sb_lstm_teacher_bwd = [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]
    sb_grad_sumsq = [_p, _p, _p, _i, _p, _p, _f, _p]
    sb_env_post = [_p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p]
    sb_loss_total = [_p, ctypes.c_double, _f, _p, _p, _p, _p, _p, _p]
    sb_gather_step = [_p, _i, _ll, _p, _ll, _p, _i, _p, _p, _p, _p, _p, _p, _p]
    sb_dense = [_p, _p, _p, _p, _i, _i, _i, _ll, _ll, _ll, _i, _i, _i, _p]
    sb_adafactor_tick = [_p, _i, _p]
    sb_adafactor_conv = [_p, _p, _i, _i, _i, _p, _f, _i, _p, _i, _p]
    sb_adafactor_small = [_p, _p, _i, _i, _p, _f, _p]
    sb_adafactor_conv_packed = [_p, _p, _p, _i, _i, _i, _p, _f, _i, _p, _i, _p]
    sb_transpose_taps = [_p, _p, _i, _i, _i, _p]
    sb_bias_act = [_p, _p, _i, _i, _p, _i, _ll, _i, _p]
    sb_transpose_taps_multi = [_p, _i, _i, _p]
    sb_adafactor_big2d = [_p, _p, _i, _i, _i, _p, _f, _p]
    # latent / heads block (csrc/latent.cu)
    sb_heads_fwd = [_p, _p, _p, _i, _p, _p, _i, _p, _p, _p, _p, _p]
    sb_heads_bwd = [_p, _p, _p, _i, ctypes.POINTER(InjGrads), _p, _p, _p, _i, _p, _p, _p, _p, _p, _p]
    sb_posterior_bits_fwd = [_p, _p, _p, _p, _p, _f, _f, _p, ctypes.c_ulonglong, _i, _p, _p, _p, _p, _p, _p]
    sb_mix_bits = [_p, _p, _p, _p, _f, _f, _p, ctypes.c_ulonglong, _i, _p, _p, _p]
    sb_posterior_bits_bwd = [_p, _p, _p, _i, _p, _p, _p]
    sb_teacher_embed_fwd = [_p, _i, _p, _p, _p, _p, _f, _p, ctypes.c_ulonglong, _p, _p, _p, _i, _p, _p]
    sb_teacher_embed_bwd = [_p, _p, _p, _f, _p, ctypes.c_ulonglong, _i, _p, _i, _p, _p, _p]
    sb_drop_scale = [_p, _p, _f, _p, ctypes.c_ulonglong, ctypes.c_uint, _p, _ll, _p]
    sb_softmax_ce_rows = [_p, _p, _i, _i, _f, _p, _p, _p, _p]
    sb_latent_out_fwd = [_p, _p, _p, _i, _p, _f, _f, _p, ctypes.c_ulonglong, _i, _p, _p, _p, _p]
    sb_latent_out_bwd = [_p, _p, _i, _p, _p, _p, _i, _p, _p, _p, _p]
    sb_sampler_tables = [_p] * 10
    sb_reward_value = [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p]
    sb_colsum_f32 = [_p, _i, _i, _i, _p, _p]


EXPORTED = ['sb_init', 'sb_launch_count', 'sb_version', 'sb_tap_wgrad_bias_ok', 'sb_set_pair_mode', 'sb_set_wgrad_tap_groups', 'sb_set_wgrad_mt2', 'sb_set_sm_reserve', 'sb_dense_splits', 'sb_set_prep_specialize', 'sb_set_ce_tma'] + [n for n in dir(_Sigs) if n.startswith('sb_')]


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f'simple_b200: {what} failed with code {rc}')


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
