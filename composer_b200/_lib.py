'''
ctypes binding of ``libcomposer_b200.so`` (the C ABI in ``include/composer_b200.h``).

The library is the product: there is no Python / CPU fallback.  If it is
missing or a call fails, an exception is raised.
'''

import ctypes
import os

_PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.path.join(_PACKAGE_DIR, 'libcomposer_b200.so')

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_u32 = ctypes.c_uint32
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p


class Config(ctypes.Structure):
    '''``cb200_config`` (include/composer_b200.h).'''

    _fields_ = [('vocab_size', ctypes.c_int32), ('embedding_size', ctypes.c_int32), ('window_size', ctypes.c_int32),
                ('decoder_layers_count', ctypes.c_int32), ('attention_head_count', ctypes.c_int32),
                ('attention_dropout_rate', c_f32), ('residual_dropout_rate', c_f32),
                ('layer_normalization_epsilon', c_f32), ('scale_attention', ctypes.c_int32),
                ('use_layer_normalization', ctypes.c_int32)]


class NativeError(RuntimeError):
    '''A call into libcomposer_b200 returned a non-zero status.'''


# name -> (restype, argtypes); restype ``c_int`` results are status codes.
_SIGNATURES = {
    'cb200_last_error': (ctypes.c_char_p, []),
    'cb200_abi_version': (c_int, []),
    'cb200_launch_count': (ctypes.c_longlong, []),
    'cb200_param_elems': (c_i64, [ctypes.POINTER(Config)]),
    'cb200_param_tensor_count': (c_int, [ctypes.POINTER(Config)]),
    'cb200_param_tensor_info': (c_int, [ctypes.POINTER(Config), c_int, ctypes.c_char_p, c_int,
                                        ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int32)]),
    'cb200_engine_create': (c_int, [ctypes.POINTER(Config), ctypes.POINTER(c_ptr)]),
    'cb200_engine_destroy': (c_int, [c_ptr]),
    'cb200_shadow_elems': (c_i64, [c_ptr]),
    'cb200_workspace_bytes': (c_i64, [c_ptr, c_int, c_int, c_int]),
    'cb200_engine_bind': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_int, c_int]),
    'cb200_refresh_shadows': (c_int, [c_ptr, c_ptr]),
    'cb200_forward': (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_u64, c_u32, c_f32, c_ptr, c_ptr, c_ptr,
                              c_ptr]),
    'cb200_backward': (c_int, [c_ptr, c_int, c_ptr]),
    'cb200_zero_grads': (c_int, [c_ptr, c_ptr]),
    'cb200_adam_step': (c_int, [c_ptr, c_f32, c_f32, c_f32, c_f32, c_i64, c_f32, c_ptr]),
    'cb200_kv_cache_elems': (c_i64, [c_ptr, c_int, c_int]),
    'cb200_decode_workspace_bytes': (c_i64, [c_ptr, c_int]),
    'cb200_generate': (c_int, [c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_f32, c_u64, c_i64,
                               c_ptr, c_ptr, c_ptr, c_ptr]),
    'cb200_prefill': (c_int, [c_ptr, c_ptr, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr]),
    'cb200_decode_step': (c_int, [c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr, c_int, c_int, c_ptr, c_ptr]),
    'cb200_set_decode_prefill': (c_int, [c_int]),
    'cb200_gemm': (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_int, c_ptr, c_int, c_ptr, c_ptr, c_int, c_ptr, c_int,
                           c_ptr, c_int, c_ptr, c_int, c_f32, c_u64, c_u32, c_u32, c_u32, c_ptr]),
    'cb200_logits_ce': (c_int, [c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32, c_ptr, c_ptr, c_ptr,
                                c_ptr]),
    'cb200_embed_fwd': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_f32, c_u64, c_u32,
                                c_ptr]),
    'cb200_embed_bwd': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_f32, c_u64, c_u32,
                                c_ptr]),
    'cb200_layernorm_fwd': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_f32, c_ptr]),
    'cb200_layernorm_bwd': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int,
                                    c_ptr]),
    'cb200_bias_grad': (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_f32, c_u64, c_u32, c_u32, c_u32, c_ptr]),
    'cb200_attention_fwd': (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_f32, c_f32, c_u64, c_u32,
                                    c_u32, c_ptr]),
    'cb200_attention_bwd': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int,
                                    c_f32, c_f32, c_u64, c_u32, c_u32, c_ptr]),
    'cb200_set_attention_fwd_impl': (c_int, [c_int]),
    'cb200_set_attention_trace': (c_int, [c_ptr]),
    'cb200_set_decode_impl': (c_int, [c_int, c_int, c_int]),
    'cb200_set_decode_profile': (c_int, [c_ptr]),
    'cb200_decode_cluster_capacity': (c_int, [c_ptr, c_int]),
    'cb200_attention_dropout_mask': (c_int, [c_ptr, c_int, c_int, c_int, c_f32, c_u64, c_u32, c_u32, c_ptr]),
    'cb200_rowmajor_dropout_mask': (c_int, [c_ptr, c_int, c_int, c_f32, c_u64, c_u32, c_u32, c_u32, c_ptr]),
    'cb200_adam': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_ptr]),
    'cb200_decode_linear': (c_int, [c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int, c_int, c_int,
                                    c_ptr]),
    'cb200_decode_attention': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_f32, c_ptr]),
}

# Entry points whose int result is a status code (0 = ok).
_STATUS = {name for name, (restype, _) in _SIGNATURES.items()
           if restype is c_int and name not in ('cb200_abi_version', 'cb200_param_tensor_count', 'cb200_decode_cluster_capacity')}

_library = None


def exported_symbols():
    '''Names every build of the library must export (what ``include/composer_b200.h`` declares).'''

    return sorted(_SIGNATURES)


def load():
    '''Loads the shared library (once) and declares the prototypes. Raises if it is missing.'''

    global _library
    if _library is not None:
        return _library

    if not os.path.exists(LIBRARY_PATH):
        raise ImportError(
            'libcomposer_b200.so is not built (%s). Run `python -m composer_b200.build`; there is no '
            'fallback path.' % LIBRARY_PATH)

    # a binary built from other sources than the ones in the tree must not pass for the product: when the hash of
    # csrc/ + include/ differs from the one recorded at build time (or none was recorded), rebuild before loading
    from composer_b200 import build as native_build
    if not native_build.is_current() and os.environ.get('CB200_ALLOW_STALE_LIBRARY') != '1':
        try:
            native_build.build()
        except Exception as error:
            raise ImportError(
                'libcomposer_b200.so does not match the sources in csrc/ and include/ (source hash differs from %s) '
                'and rebuilding it failed: %s' % (native_build.STAMP, error))

    library = ctypes.CDLL(LIBRARY_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in _SIGNATURES.items():
        function = getattr(library, name)
        function.restype = restype
        function.argtypes = argtypes

    _library = library
    return library


def call(name, *args):
    '''Calls ``name``; raises :class:`NativeError` with ``cb200_last_error`` on a non-zero status.'''

    library = load()
    result = getattr(library, name)(*args)
    if name in _STATUS and result != 0:
        message = library.cb200_last_error().decode('utf-8', 'replace')
        raise NativeError('%s failed (%d): %s' % (name, result, message))

    return result
