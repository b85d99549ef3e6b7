"""ctypes binding of libcrct_b200.so (include/crct_b200.h).  No fallback: if the library is missing or
the device is not a B200-class GPU, every call raises.

Every wrapper takes torch CUDA tensors (memory owners) and passes raw device pointers + sizes, enqueuing on
torch's current stream."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CRCT_B200_LIB') or os.path.join(_HERE, 'libcrct_b200.so')      # override: A/B runs of two builds

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_MUL, EPI_F32, EPI_BIAS_RES_F32 = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3

vp, i32, i64, f32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64


class CrctError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [('A', vp), ('B', vp), ('D', vp), ('D2', vp), ('bias', vp), ('aux', vp),
                ('M', i32), ('N', i32), ('K', i32), ('lda', i32), ('ldb', i32), ('ldd', i32), ('ldaux', i32),
                ('a_major', i32), ('b_major', i32), ('epilogue', i32), ('accumulate', i32), ('split_k', i32),
                ('block_n', i32), ('dropout_p', f32), ('seed', u64), ('max_ctas', i32), ('cta_group', i32), ('dbg', i32 * 7), ('salt', vp),
                ('a_rows_dev', vp), ('rows_hint', i32), ('drop_rows', vp)]


class LnBwdArgs(C.Structure):
    _fields_ = [('dy', vp), ('z', vp), ('mean', vp), ('rstd', vp), ('gamma', vp), ('dz', vp), ('dzm', vp),
                ('dgamma', vp), ('dbeta', vp), ('dbias', vp), ('rows', i32), ('H', i32),
                ('p_in', f32), ('seed_in', u64), ('p_out', f32), ('seed_out', u64), ('salt', vp), ('z_f32', i32), ('rows_dev', vp), ('drop_rows', vp)]


class EmbedTextArgs(C.Structure):
    _fields_ = [('ids', vp), ('types', vp), ('loc', vp), ('word', vp), ('pos', vp), ('type', vp), ('w_loc', vp),
                ('b_loc', vp), ('gamma', vp), ('beta', vp), ('y', vp), ('z', vp), ('mean', vp), ('rstd', vp),
                ('B', i32), ('T', i32), ('H', i32), ('max_pos', i32), ('dropout_p', f32), ('seed', u64), ('salt', vp),
                ('z_f32', i32), ('src_row', vp), ('rows_dev', vp), ('y32', vp)]


class EmbedTextBwdArgs(C.Structure):
    _fields_ = [('ids', vp), ('types', vp), ('loc', vp), ('dz', vp), ('g_word', vp), ('g_pos', vp), ('g_type', vp),
                ('g_wloc', vp), ('g_bloc', vp), ('B', i32), ('T', i32), ('H', i32), ('src_row', vp), ('rows_dev', vp)]


class EmbedVisArgs(C.Structure):
    _fields_ = [('g', vp), ('box', vp), ('cls', vp), ('w_loc', vp), ('b_loc', vp), ('color', vp), ('gamma', vp),
                ('beta', vp), ('y', vp), ('z', vp), ('mean', vp), ('rstd', vp), ('rows', i32), ('H', i32),
                ('dropout_p', f32), ('seed', u64), ('salt', vp), ('z_f32', i32), ('src_row', vp), ('rows_dev', vp), ('y32', vp)]


class EmbedVisBwdArgs(C.Structure):
    _fields_ = [('dz', vp), ('box', vp), ('cls', vp), ('g_color', vp), ('g_wloc', vp), ('rows', i32), ('H', i32), ('src_row', vp),
                ('rows_dev', vp)]


class AttnFwdArgs(C.Structure):
    _fields_ = [('q', vp), ('k', vp), ('v', vp), ('ldq', i32), ('ldk', i32), ('ldv', i32), ('mask_add', vp),
                ('out', vp), ('ldo', i32), ('lse', vp), ('B', i32), ('nh', i32), ('dh', i32), ('Lq', i32), ('Lk', i32),
                ('dropout_p', f32), ('seed', u64), ('salt', vp), ('cu_q', vp), ('cu_k', vp)]


class AttnBwdArgs(C.Structure):
    _fields_ = [('q', vp), ('k', vp), ('v', vp), ('ldq', i32), ('ldk', i32), ('ldv', i32), ('mask_add', vp),
                ('out', vp), ('ldo', i32), ('dout', vp), ('lddo', i32), ('lse', vp), ('dq', vp), ('dk', vp), ('dv', vp),
                ('lddq', i32), ('lddk', i32), ('lddv', i32), ('B', i32), ('nh', i32), ('dh', i32), ('Lq', i32),
                ('Lk', i32), ('dropout_p', f32), ('seed', u64), ('salt', vp), ('cu_q', vp), ('cu_k', vp)]


class LinearArgs(C.Structure):
    _fields_ = [('A', vp), ('sa_m', i64), ('sa_k', i64), ('B', vp), ('sb_k', i64), ('sb_n', i64), ('C', vp), ('ldc', i64),
                ('bias', vp), ('dmask', vp), ('ldm', i64), ('M', i32), ('N', i32), ('K', i32), ('act', i32),
                ('slope', f32), ('accumulate', i32)]


class LossArgs(C.Structure):
    _fields_ = [('logits', vp), ('reg', vp), ('labels', vp), ('R', vp), ('reg_pred', vp), ('reg_loss', vp), ('reg_l1', vp),
                ('reg_dist', vp), ('scalars', vp), ('dlogits', vp), ('dpre', vp), ('B', i32), ('l1', i32),
                ('zero_impossible', i32), ('unit_grads', i32), ('tol_margin', f32), ('nsp_coeff', f32), ('reg_coeff', f32)]


class AdamWArgs(C.Structure):
    _fields_ = [('w', vp), ('g', vp), ('m', vp), ('v', vp), ('w_bf16', vp), ('group_of_block64', vp), ('n', C.c_size_t),
                ('lr', f32 * 4), ('weight_decay', f32 * 4), ('beta1', f32), ('beta2', f32), ('eps', f32), ('step', i32),
                ('grad_scale', f32), ('dyn', vp), ('group_offset', C.c_size_t)]


class SelectArgs(C.Structure):
    _fields_ = [('logits', vp), ('reg_pred', vp), ('reg_dist', vp), ('reg_l1', vp), ('offsets', vp), ('forced', vp), ('answer', vp),
                ('prob', vp), ('sel_pred', vp), ('sel_dist', vp), ('sel_l1', vp), ('Q', i32)]


class ScoreArgs(C.Structure):
    _fields_ = [('answer', vp), ('gt_id', vp), ('needs_reg', vp), ('sel_dist', vp), ('sel_l1', vp), ('tolerance', vp), ('flags', vp),
                ('total', vp), ('Q', i32)]


MAX_CONNECTIONS = 16


class ConfigArgs(C.Structure):      # crct_config_t
    _fields_ = [('hidden_size', i32), ('num_hidden_layers', i32), ('num_attention_heads', i32), ('intermediate_size', i32),
                ('v_hidden_size', i32), ('v_num_hidden_layers', i32), ('v_num_attention_heads', i32), ('v_intermediate_size', i32),
                ('v_feature_size', i32), ('bi_hidden_size', i32), ('bi_num_attention_heads', i32), ('max_position_embeddings', i32),
                ('num_connections', i32), ('v_biattention_id', i32 * MAX_CONNECTIONS), ('t_biattention_id', i32 * MAX_CONNECTIONS),
                ('l1', i32), ('tol_margin', f32)]


class BatchArgs(C.Structure):       # crct_batch_t
    _fields_ = [('tokens', vp), ('segments', vp), ('loc', vp), ('attention_mask', vp), ('image_feat', vp), ('image_loc', vp),
                ('image_target', vp), ('image_mask', vp), ('R4', vp), ('group', vp), ('B', i32), ('Bq', i32), ('T', i32), ('R', i32),
                ('attention_mask_kind', i32), ('image_mask_kind', i32), ('text_fill', f32), ('region_fill', f32)]


class OutArgs(C.Structure):         # crct_out_t
    _fields_ = [('logits', vp), ('reg_pred', vp), ('reg_loss', vp), ('reg_l1', vp), ('reg_dist', vp), ('scalars', vp)]


# every symbol include/crct_b200.h declares (tests check the .so exports exactly these)
EXPORTS = ['crct_last_error', 'crct_version', 'crct_device_check', 'crct_gemm_bf16', 'crct_gemm_wgrad_grouped', 'crct_cast_f32_to_bf16',
           'crct_additive_mask', 'crct_layernorm_fwd', 'crct_layernorm_bwd', 'crct_layernorm_bwd_params', 'crct_colsum_bf16', 'crct_softmax_rows',
           'crct_embed_text_fwd', 'crct_embed_text_bwd', 'crct_embed_vis_fwd', 'crct_embed_vis_bwd', 'crct_attn_fwd',
           'crct_attn_bwd', 'crct_linear_f32', 'crct_linear_f32_batched', 'crct_gather_first', 'crct_scatter_first', 'crct_colsum_f32',
           'crct_pool_mul_fwd', 'crct_pool_mul_bwd', 'crct_bump_salt', 'crct_bump_salt_to', 'crct_hybrid_loss', 'crct_scale_rows', 'crct_adamw',
           'crct_expand_blocks', 'crct_select_answers', 'crct_score_answers',
           'crct_layernorm_rows_f32', 'crct_row_map', 'crct_group_map', 'crct_gather_rows', 'crct_gather_rows_f32', 'crct_scatter_rows_f32', 'crct_fill_zero',
           'crct_f32_gemm', 'crct_f32_layernorm_fwd', 'crct_f32_layernorm_bwd', 'crct_f32_layernorm_bwd_params', 'crct_f32_attn_fwd',
           'crct_f32_attn_bwd', 'crct_f32_embed_text_fwd', 'crct_f32_embed_text_bwd', 'crct_f32_embed_vis_fwd', 'crct_f32_embed_vis_bwd',
           'crct_f32_softmax_rows', 'crct_f32_gather_first', 'crct_f32_scatter_first',
           'crct_create', 'crct_destroy', 'crct_bind_params', 'crct_workspace_bytes', 'crct_forward']

_lib = None
SALT = None         # device int64[1] tensor XOR-ed into every dropout seed on the device (set by the encoder); None = off
LAUNCHES = 0          # kernels enqueued through this binding (bench.py reports it as gpu_launches)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CrctError(f'{LIB_PATH} not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                            '(there is no CPU / PyTorch fallback for the CRCT hot path)')
        _lib = C.CDLL(LIB_PATH)
        _lib.crct_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(_lib, name)          # fail loudly on a stale library
        _lib.crct_cast_f32_to_bf16.argtypes = [vp, vp, C.c_size_t, vp]
        _lib.crct_additive_mask.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        _lib.crct_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        _lib.crct_colsum_bf16.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        _lib.crct_softmax_rows.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
        _lib.crct_layernorm_rows_f32.argtypes = [vp, vp, vp, vp, C.c_longlong, vp, C.c_int, C.c_int, vp]
        _lib.crct_row_map.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        _lib.crct_group_map.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        _lib.crct_gather_rows.argtypes = [vp, vp, vp, C.c_int, C.c_longlong, vp, vp]
        _lib.crct_gather_rows_f32.argtypes = [vp, C.c_longlong, vp, vp, C.c_int, C.c_int, vp]
        _lib.crct_scatter_rows_f32.argtypes = [vp, vp, C.c_longlong, vp, C.c_int, C.c_int, vp]
        _lib.crct_fill_zero.argtypes = [vp, C.c_size_t, vp]
        _lib.crct_gather_first.argtypes = [vp, C.c_longlong, vp, C.c_int, C.c_int, vp]
        _lib.crct_scatter_first.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_int, vp]
        _lib.crct_colsum_f32.argtypes = [vp, vp, C.c_int, C.c_int, C.c_longlong, vp]
        _lib.crct_expand_blocks.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, vp]
        _lib.crct_f32_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        _lib.crct_f32_softmax_rows.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        _lib.crct_f32_gather_first.argtypes = [vp, C.c_longlong, vp, C.c_int, C.c_int, vp]
        _lib.crct_f32_scatter_first.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_int, vp]
        _lib.crct_linear_f32_batched.argtypes = [vp, C.c_int, vp]
        _lib.crct_gemm_wgrad_grouped.argtypes = [vp, C.c_int, vp]
        _lib.crct_scale_rows.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp]
        _lib.crct_pool_mul_fwd.argtypes = [vp, vp, vp, C.c_int, C.c_float, C.c_uint64, vp, vp]
        _lib.crct_pool_mul_bwd.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_uint64, vp, vp]
        _lib.crct_bump_salt.argtypes = [vp, vp]
        _lib.crct_bump_salt_to.argtypes = [vp, vp, vp]
        _lib.crct_create.argtypes = [vp, vp]
        _lib.crct_destroy.argtypes = [vp]
        _lib.crct_bind_params.argtypes = [vp, vp, vp, vp, vp, C.c_int]
        _lib.crct_workspace_bytes.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.crct_workspace_bytes.restype = C.c_size_t
        _lib.crct_forward.argtypes = [vp, vp, vp, vp, C.c_size_t, vp]
        for name in ('crct_gemm_bf16', 'crct_layernorm_bwd', 'crct_layernorm_bwd_params', 'crct_embed_text_fwd', 'crct_embed_text_bwd',
                     'crct_embed_vis_fwd', 'crct_embed_vis_bwd', 'crct_attn_fwd', 'crct_attn_bwd', 'crct_linear_f32',
                     'crct_hybrid_loss', 'crct_adamw', 'crct_select_answers', 'crct_score_answers', 'crct_f32_gemm', 'crct_f32_layernorm_bwd',
                     'crct_f32_layernorm_bwd_params', 'crct_f32_attn_fwd', 'crct_f32_attn_bwd', 'crct_f32_embed_text_fwd',
                     'crct_f32_embed_text_bwd', 'crct_f32_embed_vis_fwd', 'crct_f32_embed_vis_bwd'):
            getattr(_lib, name).argtypes = [vp, vp]
    return _lib


def check(rc: int):
    global LAUNCHES
    if rc != 0:
        raise CrctError(f'libcrct_b200 error {rc}: {lib().crct_last_error().decode()}')
    LAUNCHES += 1


def device_check():
    rc = lib().crct_device_check()
    if rc != 0:
        raise CrctError(f'libcrct_b200 error {rc}: {lib().crct_last_error().decode()}')


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def _is32(t):
    return t is not None and t.dtype == torch.float32


def _bf16(t, what):
    if t is not None and t.dtype != torch.bfloat16:
        raise CrctError(f'{what} must be bfloat16, got {t.dtype}')


def _f32(t, what):
    if t is not None and t.dtype != torch.float32:
        raise CrctError(f'{what} must be float32, got {t.dtype}')


def gemm_args(A, B, D, *, M, N, K, a_major=0, b_major=0, epilogue=EPI_BIAS, bias=None, aux=None, D2=None,
              lda=None, ldb=None, ldd=None, ldaux=None, accumulate=0, split_k=0, block_n=0, dropout_p=0.0, seed=0,
              max_ctas=0, dbg=None, cta_group=0, rows_dev=None, drop_rows=None):
    """crct_gemm_t for D[M,N] = epilogue(A·Bᵀ) — see include/crct_b200.h for operand majors (dtype checks included)."""
    f32 = _is32(A)
    if f32:
        for t, w in ((B, 'B'), (aux, 'aux'), (D2, 'D2'), (bias, 'bias'), (D, 'D')):
            _f32(t, w)
    else:
        _bf16(A, 'A'); _bf16(B, 'B'); _bf16(D2, 'D2'); _f32(bias, 'bias')
        (_f32 if epilogue == EPI_BIAS_RES_F32 else _bf16)(aux, 'aux')
        if epilogue in (EPI_F32, EPI_BIAS_RES_F32):
            _f32(D, 'D')
        else:
            _bf16(D, 'D')
    a = GemmArgs()
    a.A, a.B, a.D, a.D2, a.bias, a.aux = ptr(A), ptr(B), ptr(D), ptr(D2), ptr(bias), ptr(aux)
    a.M, a.N, a.K = M, N, K
    a.lda = lda if lda is not None else A.stride(0)
    a.ldb = ldb if ldb is not None else B.stride(0)
    a.ldd = ldd if ldd is not None else D.stride(0)
    a.ldaux = ldaux if ldaux is not None else (aux.stride(0) if aux is not None else 0)
    a.a_major, a.b_major, a.epilogue = a_major, b_major, epilogue
    a.accumulate, a.split_k, a.block_n = accumulate, split_k, block_n
    a.dropout_p, a.seed, a.max_ctas, a.cta_group = dropout_p, seed, max_ctas, cta_group
    a.salt = ptr(SALT)
    a.a_rows_dev, a.drop_rows = ptr(rows_dev), ptr(drop_rows)
    a.rows_hint = int(getattr(rows_dev, 'hint', 0) or 0)      # expected row count (tile-shape choice only)
    if f32 and rows_dev is not None:
        raise CrctError('the fp32 check mode runs the padded layout (no device-side row counts)')
    if dbg is not None:
        for i, v in enumerate(dbg):
            a.dbg[i] = v
    return a


def gemm(A, B, D, **kw):
    """D[M,N] = epilogue(A·Bᵀ).  fp32 operands select the check-mode kernel."""
    a = gemm_args(A, B, D, **kw)
    check((lib().crct_f32_gemm if _is32(A) else lib().crct_gemm_bf16)(C.byref(a), stream_ptr()))


def gemm_wgrad_grouped(problems):
    """Up to 8 weight-gradient problems (gemm_args(..., a_major=1, b_major=1, epilogue=EPI_F32, accumulate=1)) in one launch."""
    arr = (GemmArgs * len(problems))(*problems)
    check(lib().crct_gemm_wgrad_grouped(arr, len(problems), stream_ptr()))


def cast_f32_to_bf16(src, dst):
    _f32(src, 'src'); _bf16(dst, 'dst')
    check(lib().crct_cast_f32_to_bf16(ptr(src), ptr(dst), src.numel(), stream_ptr()))


def additive_mask(mask, out):
    kind = {torch.bool: 0, torch.uint8: 0, torch.int64: 1, torch.float32: 2}.get(mask.dtype)
    if kind is None:
        raise CrctError(f'unsupported mask dtype {mask.dtype}')
    check(lib().crct_additive_mask(ptr(mask), kind, ptr(out), mask.numel(), stream_ptr()))


def layernorm_fwd(z, gamma, beta, y, mean=None, rstd=None, rows_dev=None, y32=None):
    """y bf16 <- LayerNorm(z), z bf16 or fp32 (production keeps the pre-LayerNorm sum in fp32); fp32 y selects the check mode."""
    rows, H = z.shape
    if _is32(y):
        _f32(z, 'z'); _f32(gamma, 'gamma')
        return check(lib().crct_f32_layernorm_fwd(ptr(z), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, H, stream_ptr()))
    _bf16(y, 'y'); _f32(gamma, 'gamma')
    if not _is32(z):
        _bf16(z, 'z')
    _f32(y32, 'y32')
    check(lib().crct_layernorm_fwd(ptr(z), ptr(gamma), ptr(beta), ptr(y), ptr(y32), ptr(mean), ptr(rstd), rows, H, int(_is32(z)), ptr(rows_dev),
                                   stream_ptr()))


def _ln_bwd_args(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dbias, dzm, p_in, seed_in, p_out, seed_out, rows_dev=None, drop_rows=None):
    chk = _f32 if _is32(dy) else _bf16
    chk(dy, 'dy'); chk(dz, 'dz'); chk(dzm, 'dzm')
    if not _is32(z):
        chk(z, 'z')
    a = LnBwdArgs()
    a.dy, a.z, a.mean, a.rstd, a.gamma, a.dz, a.dzm = ptr(dy), ptr(z), ptr(mean), ptr(rstd), ptr(gamma), ptr(dz), ptr(dzm)
    a.dgamma, a.dbeta, a.dbias = ptr(dgamma), ptr(dbeta), ptr(dbias)
    a.rows, a.H = z.shape
    a.p_in, a.seed_in, a.p_out, a.seed_out = p_in, seed_in, p_out, seed_out
    a.salt = ptr(SALT)
    a.z_f32, a.rows_dev, a.drop_rows = int(_is32(z) and not _is32(dy)), ptr(rows_dev), ptr(drop_rows)
    return a


def layernorm_bwd(dy, z, mean, rstd, gamma, dz, dgamma=None, dbeta=None, dbias=None, dzm=None, p_in=0.0, seed_in=0, p_out=0.0,
                  seed_out=0, rows_dev=None, drop_rows=None):
    """dgamma = dbeta = dbias = None: input gradient only (see layernorm_bwd_params)."""
    a = _ln_bwd_args(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dbias, dzm, p_in, seed_in, p_out, seed_out, rows_dev, drop_rows)
    check((lib().crct_f32_layernorm_bwd if _is32(dy) else lib().crct_layernorm_bwd)(C.byref(a), stream_ptr()))


def layernorm_bwd_params(dy, z, mean, rstd, dz, dgamma, dbeta, dbias=None, dzm=None, p_in=0.0, seed_in=0, p_out=0.0, rows_dev=None,
                         drop_rows=None):
    """Column sums of the split LayerNorm backward: dgamma, dbeta and (from dzm, or dz when p_out == 0) dbias."""
    a = _ln_bwd_args(dy, z, mean, rstd, None, dz, dgamma, dbeta, dbias, dzm, p_in, seed_in, p_out, 0, rows_dev, drop_rows)
    check((lib().crct_f32_layernorm_bwd_params if _is32(dy) else lib().crct_layernorm_bwd_params)(C.byref(a), stream_ptr()))


def colsum_bf16(x, out, rows=None, N=None, ld=None, rows_dev=None):
    rows = x.shape[0] if rows is None else rows
    N = x.shape[1] if N is None else N
    if _is32(x):
        return colsum_f32(x, out, rows, N, x.stride(0) if ld is None else ld)
    _bf16(x, 'x'); _f32(out, 'out')
    check(lib().crct_colsum_bf16(ptr(x), ptr(out), rows, N, x.stride(0) if ld is None else ld, ptr(rows_dev), stream_ptr()))


def softmax_rows(x, out, src_row=None, rows_dev=None):
    """out[r] = softmax(x[src_row[r]]) (src_row None: r); out may have fewer (packed) rows than x."""
    rows, F = out.shape
    if _is32(out):
        return check(lib().crct_f32_softmax_rows(ptr(x), ptr(out), rows, F, stream_ptr()))
    _f32(x, 'x'); _bf16(out, 'out')
    check(lib().crct_softmax_rows(ptr(x), ptr(out), rows, F, ptr(src_row), ptr(rows_dev), stream_ptr()))


def embed_text_fwd(ids, types, loc, word, pos, type_, w_loc, b_loc, gamma, beta, y, z=None, mean=None, rstd=None,
                   dropout_p=0.0, seed=0, src_row=None, rows_dev=None, y32=None):
    a = EmbedTextArgs()
    a.ids, a.types, a.loc = ptr(ids), ptr(types), ptr(loc)
    a.word, a.pos, a.type, a.w_loc, a.b_loc, a.gamma, a.beta = (ptr(word), ptr(pos), ptr(type_), ptr(w_loc), ptr(b_loc),
                                                                ptr(gamma), ptr(beta))
    a.y, a.z, a.mean, a.rstd = ptr(y), ptr(z), ptr(mean), ptr(rstd)
    a.B, a.T = ids.shape
    a.H, a.max_pos = word.shape[1], pos.shape[0]
    a.dropout_p, a.seed, a.salt = dropout_p, seed, ptr(SALT)
    a.z_f32, a.src_row, a.rows_dev, a.y32 = int(_is32(z) and not _is32(y)), ptr(src_row), ptr(rows_dev), ptr(y32)
    check((lib().crct_f32_embed_text_fwd if _is32(y) else lib().crct_embed_text_fwd)(C.byref(a), stream_ptr()))


def embed_text_bwd(ids, types, loc, dz, g_word, g_pos, g_type, g_wloc, g_bloc, src_row=None, rows_dev=None):
    a = EmbedTextBwdArgs()
    a.ids, a.types, a.loc, a.dz = ptr(ids), ptr(types), ptr(loc), ptr(dz)
    a.g_word, a.g_pos, a.g_type, a.g_wloc, a.g_bloc = ptr(g_word), ptr(g_pos), ptr(g_type), ptr(g_wloc), ptr(g_bloc)
    a.B, a.T = ids.shape
    a.H = dz.shape[-1]
    a.src_row, a.rows_dev = ptr(src_row), ptr(rows_dev)
    check((lib().crct_f32_embed_text_bwd if _is32(dz) else lib().crct_embed_text_bwd)(C.byref(a), stream_ptr()))


def embed_vis_fwd(g, box, cls, w_loc, b_loc, color, gamma, beta, y, z=None, mean=None, rstd=None, dropout_p=0.0, seed=0,
                  src_row=None, rows_dev=None, y32=None):
    a = EmbedVisArgs()
    a.g, a.box, a.cls, a.w_loc, a.b_loc, a.color, a.gamma, a.beta = (ptr(g), ptr(box), ptr(cls), ptr(w_loc), ptr(b_loc),
                                                                     ptr(color), ptr(gamma), ptr(beta))
    a.y, a.z, a.mean, a.rstd = ptr(y), ptr(z), ptr(mean), ptr(rstd)
    a.rows, a.H = g.shape
    a.dropout_p, a.seed, a.salt = dropout_p, seed, ptr(SALT)
    a.z_f32, a.src_row, a.rows_dev, a.y32 = int(_is32(z) and not _is32(y)), ptr(src_row), ptr(rows_dev), ptr(y32)
    check((lib().crct_f32_embed_vis_fwd if _is32(y) else lib().crct_embed_vis_fwd)(C.byref(a), stream_ptr()))


def embed_vis_bwd(dz, box, cls, g_color, g_wloc, src_row=None, rows_dev=None):
    a = EmbedVisBwdArgs()
    a.dz, a.box, a.cls, a.g_color, a.g_wloc = ptr(dz), ptr(box), ptr(cls), ptr(g_color), ptr(g_wloc)
    a.rows, a.H = dz.shape
    a.src_row, a.rows_dev = ptr(src_row), ptr(rows_dev)
    check((lib().crct_f32_embed_vis_bwd if _is32(dz) else lib().crct_embed_vis_bwd)(C.byref(a), stream_ptr()))


def attn_fwd(q, k, v, mask_add, out, lse, *, B, nh, dh, Lq, Lk, ldq, ldk, ldv, ldo, dropout_p=0.0, seed=0, cu_q=None, cu_k=None):
    a = AttnFwdArgs()
    a.q, a.k, a.v, a.mask_add, a.out, a.lse = ptr(q), ptr(k), ptr(v), ptr(mask_add), ptr(out), ptr(lse)
    a.ldq, a.ldk, a.ldv, a.ldo = ldq, ldk, ldv, ldo
    a.B, a.nh, a.dh, a.Lq, a.Lk = B, nh, dh, Lq, Lk
    a.dropout_p, a.seed, a.salt = dropout_p, seed, ptr(SALT)
    a.cu_q, a.cu_k = ptr(cu_q), ptr(cu_k)
    check((lib().crct_f32_attn_fwd if _is32(q) else lib().crct_attn_fwd)(C.byref(a), stream_ptr()))


def attn_bwd(q, k, v, mask_add, out, dout, lse, dq, dk, dv, *, B, nh, dh, Lq, Lk, ldq, ldk, ldv, ldo, lddo, lddq, lddk,
             lddv, dropout_p=0.0, seed=0, cu_q=None, cu_k=None):
    a = AttnBwdArgs()
    a.q, a.k, a.v, a.mask_add, a.out, a.dout, a.lse = ptr(q), ptr(k), ptr(v), ptr(mask_add), ptr(out), ptr(dout), ptr(lse)
    a.dq, a.dk, a.dv = ptr(dq), ptr(dk), ptr(dv)
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddo, a.lddq, a.lddk, a.lddv = ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv
    a.B, a.nh, a.dh, a.Lq, a.Lk = B, nh, dh, Lq, Lk
    a.dropout_p, a.seed, a.salt = dropout_p, seed, ptr(SALT)
    a.cu_q, a.cu_k = ptr(cu_q), ptr(cu_k)
    check((lib().crct_f32_attn_bwd if _is32(q) else lib().crct_attn_bwd)(C.byref(a), stream_ptr()))


def linear_f32(A, sa_m, sa_k, B, sb_k, sb_n, Cout, ldc, M, N, K, bias=None, act=ACT_NONE, dmask=None, ldm=0, slope=0.0,
               accumulate=0):
    a = LinearArgs()
    a.A, a.sa_m, a.sa_k, a.B, a.sb_k, a.sb_n, a.C, a.ldc = ptr(A), sa_m, sa_k, ptr(B), sb_k, sb_n, ptr(Cout), ldc
    a.bias, a.dmask, a.ldm = ptr(bias), ptr(dmask), ldm
    a.M, a.N, a.K, a.act, a.slope, a.accumulate = M, N, K, act, slope, accumulate
    check(lib().crct_linear_f32(C.byref(a), stream_ptr()))


def lin_problem(A, sa_m, sa_k, B, sb_k, sb_n, Cout, ldc, M, N, K, bias=None, act=ACT_NONE, dmask=None, ldm=0, slope=0.0, accumulate=0):
    a = LinearArgs()
    a.A, a.sa_m, a.sa_k, a.B, a.sb_k, a.sb_n, a.C, a.ldc = ptr(A), sa_m, sa_k, ptr(B), sb_k, sb_n, ptr(Cout), ldc
    a.bias, a.dmask, a.ldm = ptr(bias), ptr(dmask), ldm
    a.M, a.N, a.K, a.act, a.slope, a.accumulate = M, N, K, act, slope, accumulate
    return a


def linear_f32_batched(problems):
    """Independent small fp32 problems in one launch (include/crct_b200.h: crct_linear_f32_batched)."""
    arr = (LinearArgs * len(problems))(*problems)
    check(lib().crct_linear_f32_batched(arr, len(problems), stream_ptr()))


def gather_first(src, row_stride, out):
    B, H = out.shape
    check((lib().crct_f32_gather_first if _is32(src) else lib().crct_gather_first)(ptr(src), row_stride, ptr(out), B, H, stream_ptr()))


def scatter_first(g, dst, row_stride):
    B, H = g.shape
    check((lib().crct_f32_scatter_first if _is32(dst) else lib().crct_scatter_first)(ptr(g), ptr(dst), row_stride, B, H, stream_ptr()))


def colsum_f32(x, out, M, N, ld):
    check(lib().crct_colsum_f32(ptr(x), ptr(out), M, N, ld, stream_ptr()))


def pool_mul_fwd(pt, pv, out, p=0.0, seed=0):
    check(lib().crct_pool_mul_fwd(ptr(pt), ptr(pv), ptr(out), pt.numel(), p, seed, ptr(SALT), stream_ptr()))


def pool_mul_bwd(dpooled, pt, pv, dut, duv, p=0.0, seed=0):
    check(lib().crct_pool_mul_bwd(ptr(dpooled), ptr(pt), ptr(pv), ptr(dut), ptr(duv), pt.numel(), p, seed, ptr(SALT), stream_ptr()))


def hybrid_loss(logits, reg, labels, R, reg_pred, reg_loss, reg_l1, reg_dist, scalars, dlogits=None, dpre=None, *, l1,
                zero_impossible, tol_margin, nsp_coeff=1.0, reg_coeff=1.0, unit_grads=0):
    a = LossArgs()
    a.logits, a.reg, a.labels, a.R = ptr(logits), ptr(reg), ptr(labels), ptr(R)
    a.reg_pred, a.reg_loss, a.reg_l1, a.reg_dist, a.scalars = ptr(reg_pred), ptr(reg_loss), ptr(reg_l1), ptr(reg_dist), ptr(scalars)
    a.dlogits, a.dpre = ptr(dlogits), ptr(dpre)
    a.B, a.l1, a.zero_impossible, a.unit_grads = logits.shape[0], int(l1), int(zero_impossible), int(unit_grads)
    a.tol_margin, a.nsp_coeff, a.reg_coeff = tol_margin, nsp_coeff, reg_coeff
    check(lib().crct_hybrid_loss(C.byref(a), stream_ptr()))


def scale_rows(x, s, out):
    B, n = x.shape
    check(lib().crct_scale_rows(ptr(x), ptr(s), 1 if s.numel() == B else 0, ptr(out), B, n, stream_ptr()))


def bump_salt(salt, snapshot=None):
    if snapshot is None:
        return check(lib().crct_bump_salt(ptr(salt), stream_ptr()))
    check(lib().crct_bump_salt_to(ptr(salt), ptr(snapshot), stream_ptr()))


def adamw(w, g, m, v, w_bf16, group, n, lr4, wd4, beta1, beta2, eps, step, grad_scale=1.0, dyn=None, group_offset=0):
    a = AdamWArgs()
    a.group_offset = group_offset
    a.w, a.g, a.m, a.v, a.w_bf16, a.group_of_block64, a.n = ptr(w), ptr(g), ptr(m), ptr(v), ptr(w_bf16), ptr(group), n
    for i in range(4):
        a.lr[i], a.weight_decay[i] = lr4[i], wd4[i]
    a.beta1, a.beta2, a.eps, a.step, a.grad_scale, a.dyn = beta1, beta2, eps, step, grad_scale, ptr(dyn)
    check(lib().crct_adamw(C.byref(a), stream_ptr()))


def expand_blocks(src, group, dst):
    """dst[n] = src[group[n]] over the leading dimension (include/crct_b200.h: crct_expand_blocks)."""
    n = group.numel()
    if n == 0:
        return
    per = dst.numel() // max(n, 1) * dst.element_size()
    assert src.dtype == dst.dtype and group.dtype == torch.int64 and src.is_contiguous() and dst.is_contiguous()
    check(lib().crct_expand_blocks(ptr(src), ptr(group), ptr(dst), n, per, stream_ptr()))


def select_answers(logits, reg_pred, reg_dist, reg_l1, offsets, answer, sel_pred, sel_dist, sel_l1, prob=None, forced=None):
    a = SelectArgs()
    a.logits, a.reg_pred, a.reg_dist, a.reg_l1 = ptr(logits), ptr(reg_pred), ptr(reg_dist), ptr(reg_l1)
    a.offsets, a.forced, a.answer, a.prob = ptr(offsets), ptr(forced), ptr(answer), ptr(prob)
    a.sel_pred, a.sel_dist, a.sel_l1, a.Q = ptr(sel_pred), ptr(sel_dist), ptr(sel_l1), answer.numel()
    check(lib().crct_select_answers(C.byref(a), stream_ptr()))


def score_answers(answer, gt_id, needs_reg, sel_dist, sel_l1, tolerance, total, flags=None):
    a = ScoreArgs()
    a.answer, a.gt_id, a.needs_reg, a.sel_dist, a.sel_l1 = ptr(answer), ptr(gt_id), ptr(needs_reg), ptr(sel_dist), ptr(sel_l1)
    a.tolerance, a.flags, a.total, a.Q = ptr(tolerance), ptr(flags), ptr(total), answer.numel()
    check(lib().crct_score_answers(C.byref(a), stream_ptr()))


# ---- var-len ("packed") rows: include/crct_b200.h "Var-len rows"
def row_map(mask, cu, src_row):
    kind = {torch.bool: 0, torch.uint8: 0, torch.int64: 1, torch.float32: 2}.get(mask.dtype)
    if kind is None:
        raise CrctError(f'unsupported mask dtype {mask.dtype}')
    B, Lm = mask.shape
    assert cu.dtype == torch.int32 and src_row.dtype == torch.int32 and cu.numel() >= B + 1 and src_row.numel() >= B * Lm
    check(lib().crct_row_map(ptr(mask), kind, B, Lm, ptr(cu), ptr(src_row), stream_ptr()))


def group_map(src_cu, group, cu, src_row):
    assert group.dtype == torch.int64 and cu.dtype == torch.int32 and src_row.dtype == torch.int32
    check(lib().crct_group_map(ptr(src_cu), ptr(group), group.numel(), ptr(cu), ptr(src_row), stream_ptr()))


def gather_rows(src, idx, dst, rows_dev=None):
    """dst[r] = src[idx[r]] over the leading dimension (16-byte multiples per row)."""
    assert src.dtype == dst.dtype and idx.dtype == torch.int32 and src.is_contiguous() and dst.is_contiguous()
    rows = dst.shape[0]
    per = dst.numel() // max(rows, 1) * dst.element_size()
    check(lib().crct_gather_rows(ptr(src), ptr(idx), ptr(dst), rows, per, ptr(rows_dev), stream_ptr()))


def gather_rows_f32(src, row_index, out):
    B, H = out.shape
    _bf16(src, 'src'); _f32(out, 'out')
    check(lib().crct_gather_rows_f32(ptr(src), src.stride(0), ptr(row_index), ptr(out), B, H, stream_ptr()))


def scatter_rows_f32(g, dst, row_index):
    B, H = g.shape
    _bf16(dst, 'dst'); _f32(g, 'g')
    check(lib().crct_scatter_rows_f32(ptr(g), ptr(dst), dst.stride(0), ptr(row_index), B, H, stream_ptr()))


def fill_zero(t):
    assert t.is_contiguous()
    check(lib().crct_fill_zero(ptr(t), t.numel() * t.element_size(), stream_ptr()))


def layernorm_rows_f32(z, gamma, beta, out, row_index=None, row_step=0):
    """out[b] = LayerNorm(z[row_index[b]]) (or z[b * row_step]) in fp32 — the heads' inputs."""
    B, H = out.shape
    _f32(z, 'z'); _f32(out, 'out')
    check(lib().crct_layernorm_rows_f32(ptr(z), ptr(gamma), ptr(beta), ptr(row_index), row_step, ptr(out), B, H, stream_ptr()))
