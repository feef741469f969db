"""troy-nova_b200 — python (ctypes) binding of libb2rns.so, the B200-native RNS polynomial hot path.

The product is the C-ABI library (include/b2rns.h) plus the C++ host mirror of the reference's
Evaluator seam (troy-nova_b200/host/).  This module exists so that tests/ and bench.py can drive
the C ABI with torch-allocated device memory; it contains no arithmetic and no fallback: if the
CUDA library is missing or no GPU is present every compute call raises.

Import with  importlib.import_module("troy-nova_b200")  (the directory name is not an identifier).
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2RNS_LIB") or os.path.join(HERE, "libb2rns.so")   # B2RNS_LIB: developer override for A/B builds

BFV, CKKS, BGV = 1, 2, 3
ASSIGN_OVERWRITE, ASSIGN_ADD_INPLACE, ASSIGN_OVERWRITE_EXCEPT_FIRST = 0, 1, 2
TABLES_COMPONENTWISE, TABLES_KS_SET_PRODUCTS, TABLES_KS_SKIP_FINALS = 0, 1, 2

u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class B2Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("b2 status %d: %s" % (status, msg))
        self.status = status


def build(verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libb2rns.so (in-tree)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE], stdout=None if verbose else subprocess.DEVNULL)


_lib = None

_SIGS = {
    # name: (restype, argtypes)
    "b2_last_error": (C.c_char_p, []),
    "b2_version": (C.c_char_p, []),
    "b2_device_count": (C.c_int, []),
    "b2_launch_count": (C.c_ulonglong, []),
    "b2_ctx_destroy": (None, [vp]),
    "b2_coeff_modulus_create": (C.c_int, [C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t, u64p]),
    "b2_ctx_root": (C.c_int, [vp, C.c_size_t, u64p]),
    "b2_malloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "b2_free": (C.c_int, [vp, vp]),
    "b2_malloc_host": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "b2_free_host": (C.c_int, [vp]),
    "b2_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "b2_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "b2_memcpy_d2d": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "b2_stream_sync": (C.c_int, [vp, vp]),
}
_NTT_ARGS = [vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, vp]
for _n in ("b2_ntt_forward", "b2_ntt_inverse", "b2_ntt_forward_host", "b2_ntt_inverse_host"):
    _SIGS[_n] = (C.c_int, _NTT_ARGS)
for _n in ("b2_ntt_forward_ptrs", "b2_ntt_inverse_ptrs"):
    _SIGS[_n] = (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, vp])


_SIGS.update({
    "b2_kswitch_key_create": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp)]),
    "b2_kswitch_key_create_contiguous": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(vp)]),
    "b2_kswitch_key_wrap": (C.c_int, [vp, vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp)]),
    "b2_kswitch_key_destroy": (None, [vp]),
    "b2_keyswitch": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, vp, C.c_int, vp]),
    "b2_keyswitch_host": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, vp, C.c_int, vp]),
    "b2_keyswitch_batched": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_int, vp, C.c_int, vp]),
    "b2_limb_partition": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "b2_limb_partition_weighted": (C.c_int, [vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "b2_shard_push": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_size_t, C.POINTER(vp), C.c_size_t, C.c_uint64, vp]),
    "b2_shard_wait": (C.c_int, [vp, vp, C.POINTER(C.c_int), C.c_int, C.c_uint64, vp]),
    "b2_kswitch_key_shard": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(vp)]),
    "b2_keyswitch_shard_prepare": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, vp, vp]),
    "b2_keyswitch_shard_products": (C.c_int, [vp, vp, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp, vp]),
    "b2_keyswitch_shard_finish": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, vp, C.c_int, vp]),
    "b2_mod_switch_to_next": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp]),
    "b2_add": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp]),
    "b2_sub": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp]),
    "b2_negate": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp]),
    "b2_dyadic_convolute": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, vp]),
    "b2_dyadic_square": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_size_t, vp]),
    "b2_dyadic_broadcast": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp]),
    "b2_dyadic_broadcast_batched": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp]),
    "b2_multiply_plain_accumulate": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp]),
    "b2_negacyclic_shift": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp]),
    "b2_galois": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp]),
    "b2_bfv_multiply": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_size_t, vp]),
    "b2_behz_bases": (C.c_int, [vp, C.c_size_t, u64p, C.POINTER(C.c_size_t), u64p]),
    "b2_scale_inv_degree": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, vp]),
    "b2_prng_fill": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "b2_prng_sample_uint64": (C.c_uint64, [vp]),
    "b2_sample_poly": (C.c_int, [vp, C.c_int, vp, vp, C.c_size_t, C.c_size_t, vp]),
    "b2_encrypt_zero_asymmetric": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, vp, vp, vp]),
    "b2_encrypt_zero_symmetric": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.c_int, vp, vp, vp, u64p, vp]),
    "b2_bfv_scale_up": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, vp, C.c_int, vp, vp]),
    "b2_centralize": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, vp, vp]),
    "b2_centralize_scaled": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, C.c_uint64, vp, vp]),
    "b2_encrypt": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_int, C.c_size_t, vp, vp, vp, vp]),
    "b2_pack_layer": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, vp, vp]),
    "b2_pack_rlwes": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, u64p, C.POINTER(vp), C.c_size_t,
                                C.c_int, C.c_int, C.POINTER(vp), vp]),
    "b2_extract_lwe": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, vp, vp, vp]),
    "b2_assemble_lwes": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, vp]),
    "b2_pack_lwes": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t, u64p, C.POINTER(vp), C.c_size_t, C.c_int, C.c_int, C.POINTER(vp), vp]),
    "b2_batch_encode": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
    "b2_batch_decode": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
    "b2_generate_secret_key": (C.c_int, [vp, vp, vp, vp]),
    "b2_generate_kswitch_key": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "b2_secret_key_powers": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
    "b2_decrypt_phase": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp, vp]),
    "b2_decrypt_scale_and_round": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, vp]),
    "b2_decrypt_mod_t": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, u64p, vp]),
    "b2_decrypt": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_size_t, vp, u64p, vp]),
})
MODSWITCH_BFV, MODSWITCH_CKKS, MODSWITCH_BGV, MODSWITCH_DROP = 0, 1, 2, 3
SAMPLE_TERNARY, SAMPLE_CENTERED_BINOMIAL, SAMPLE_UNIFORM = 1, 2, 3


class Prng(C.Structure):
    """utils::RandomGenerator state (src/utils/random_generator.h:45-93): AES-128-CTR, advanced in place by every call that draws from it"""
    _fields_ = [("seed_lo", C.c_uint64), ("seed_hi", C.c_uint64), ("counter_lo", C.c_uint64), ("counter_hi", C.c_uint64)]

    def __init__(self, seed_lo=0, seed_hi=0, counter=0):
        super().__init__(seed_lo, seed_hi, counter & 0xFFFFFFFFFFFFFFFF, counter >> 64)

    def state(self):
        return (self.seed_lo, self.seed_hi, self.counter_lo, self.counter_hi)


class _Params(C.Structure):
    _fields_ = [("log_n", C.c_uint32), ("n_moduli", C.c_uint32), ("moduli", u64p), ("plain_modulus", C.c_uint64), ("scheme", C.c_uint32)]


def lib():
    """Load libb2rns.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libb2rns.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C troy-nova_b200`")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L.b2_ctx_create.restype = C.c_int
        L.b2_ctx_create.argtypes = [C.c_int, C.POINTER(_Params), C.POINTER(vp)]
        _lib = L
    return _lib


def exported_symbols():
    return sorted(list(_SIGS) + ["b2_ctx_create"])


def _chk(st):
    if st != 0:
        raise B2Error(st, lib().b2_last_error().decode())


def launch_count():
    return int(lib().b2_launch_count())


def coeff_modulus_create(n, bits):
    out = (C.c_uint64 * len(bits))()
    _chk(lib().b2_coeff_modulus_create(n, (C.c_uint32 * len(bits))(*bits), len(bits), out))
    return [int(x) for x in out]


def limb_partition(n_out_moduli, world, rank):
    """(first, count) of the output moduli 0..n_out_moduli-1 owned by `rank` (host-only; b2_limb_partition)."""
    f, c = C.c_size_t(), C.c_size_t()
    _chk(lib().b2_limb_partition(n_out_moduli, world, rank, C.byref(f), C.byref(c)))
    return int(f.value), int(c.value)


def _dptr(x):
    """device pointer of a torch tensor / int / None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    return x.data_ptr()


class Context:
    """Device-resident tables for one modulus chain (key level, special prime last)."""

    def __init__(self, log_n, moduli, plain_modulus=0, scheme=CKKS, device=0):
        self.L = lib()
        self.log_n, self.n = log_n, 1 << log_n
        self.moduli = [int(m) for m in moduli]
        self.K = len(self.moduli)
        self.t, self.scheme, self.device = int(plain_modulus), scheme, device
        arr = (C.c_uint64 * self.K)(*self.moduli)
        prm = _Params(log_n, self.K, arr, self.t, scheme)
        h = vp()
        _chk(self.L.b2_ctx_create(device, C.byref(prm), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.b2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def root(self, i):
        v = C.c_uint64()
        _chk(self.L.b2_ctx_root(self.h, i, C.byref(v)))
        return int(v.value)

    @staticmethod
    def _stream(stream):
        if stream is None:
            import torch
            return torch.cuda.current_stream().cuda_stream
        return stream

    # N1 / N2 ---------------------------------------------------------------------------------
    def ntt_forward(self, src, dst=None, batch=1, pcount=1, ncomp=1, mode=TABLES_COMPONENTWISE, table_base=0, decomp=0, stream=None):
        dst = src if dst is None else dst
        _chk(self.L.b2_ntt_forward(self.h, _dptr(src), _dptr(dst), batch, pcount, ncomp, mode, table_base, decomp, self._stream(stream)))
        return dst

    def ntt_inverse(self, src, dst=None, batch=1, pcount=1, ncomp=1, mode=TABLES_COMPONENTWISE, table_base=0, decomp=0, stream=None):
        dst = src if dst is None else dst
        _chk(self.L.b2_ntt_inverse(self.h, _dptr(src), _dptr(dst), batch, pcount, ncomp, mode, table_base, decomp, self._stream(stream)))
        return dst

    def ntt_ptrs(self, forward, srcs, dsts=None, pcount=1, ncomp=1, mode=TABLES_COMPONENTWISE, table_base=0, decomp=0, stream=None):
        """scattered batch (the reference's *_batched over a SliceVec): lists of device tensors / raw pointers, ONE launch"""
        dsts = srcs if dsts is None else dsts
        nb = len(srcs)
        ip = (vp * nb)(*[_dptr(t) for t in srcs]); op = (vp * nb)(*[_dptr(t) for t in dsts])
        f = self.L.b2_ntt_forward_ptrs if forward else self.L.b2_ntt_inverse_ptrs
        _chk(f(self.h, ip, op, nb, pcount, ncomp, mode, table_base, decomp, self._stream(stream)))
        return dsts

    def ntt_host(self, forward, src_np, dst_np, batch=1, pcount=1, ncomp=1, mode=TABLES_COMPONENTWISE, table_base=0, decomp=0, stream=0):
        f = self.L.b2_ntt_forward_host if forward else self.L.b2_ntt_inverse_host
        _chk(f(self.h, src_np.ctypes.data, dst_np.ctypes.data, batch, pcount, ncomp, mode, table_base, decomp, stream))
        return dst_np

    # K1-K5 -------------------------------------------------------------------------------------
    def kswitch_key(self, key_tensor, digits=None):
        """key_tensor: device tensor holding [digits][2][K][N] (kept alive by the returned handle)"""
        digits = self.K - 1 if digits is None else digits
        h = vp()
        _chk(self.L.b2_kswitch_key_create_contiguous(self.h, _dptr(key_tensor), digits, C.byref(h)))
        return KSwitchKey(self, h, key_tensor)

    def keyswitch(self, target, decomp, is_ntt, key, dest, assign=ASSIGN_OVERWRITE, stream=None):
        _chk(self.L.b2_keyswitch(self.h, _dptr(target), decomp, int(is_ntt), key.h, _dptr(dest), assign, self._stream(stream)))
        return dest

    def keyswitch_host(self, target_np, decomp, is_ntt, key, dest_np, assign=ASSIGN_OVERWRITE, stream=0):
        """target / dest in host memory (numpy uint64, ideally pinned); key resident on the device; synchronised on return"""
        _chk(self.L.b2_keyswitch_host(self.h, target_np.ctypes.data, decomp, int(is_ntt), key.h, dest_np.ctypes.data, assign, stream))
        return dest_np

    def keyswitch_batched(self, targets, decomp, is_ntt, key, dests, assign=ASSIGN_OVERWRITE, stream=None):
        """targets / dests: lists of device tensors (or raw device pointers), one per ciphertext"""
        nb = len(targets)
        tp = (vp * nb)(*[_dptr(t) for t in targets]); dp = (vp * nb)(*[_dptr(d) for d in dests])
        _chk(self.L.b2_keyswitch_batched(self.h, tp, dp, nb, decomp, int(is_ntt), key.h, assign, self._stream(stream)))
        return dests

    # limb-sharded keyswitch stages (multi-GPU, see troy-nova_b200/sharded.py) -------------------
    def kswitch_key_shard(self, full_key, decomp, first, count):
        h = vp()
        _chk(self.L.b2_kswitch_key_shard(self.h, full_key.h, decomp, first, count, C.byref(h)))
        return KSwitchKey(self, h, None)

    def limb_partition_weighted(self, decomp, world, rank):
        f, c = C.c_size_t(), C.c_size_t()
        _chk(self.L.b2_limb_partition_weighted(self.h, decomp, world, rank, C.byref(f), C.byref(c)))
        return int(f.value), int(c.value)

    def shard_push(self, src, peer_dst, words, peer_flags, flag_slot, epoch, stream=None):
        """copy `words` u64 from src into every peer's buffer (raw device pointers), then raise flag `flag_slot` to `epoch` on every peer"""
        w = len(peer_dst)
        dp = (vp * w)(*[_dptr(p) for p in peer_dst]); fp = (vp * w)(*[_dptr(p) for p in peer_flags])
        _chk(self.L.b2_shard_push(self.h, _dptr(src), dp, w, words, fp, flag_slot, epoch, self._stream(stream)))

    def shard_wait(self, flags, slots, epoch, stream=None):
        arr = (C.c_int * len(slots))(*slots)
        _chk(self.L.b2_shard_wait(self.h, _dptr(flags), arr, len(slots), epoch, self._stream(stream)))

    def keyswitch_shard_prepare(self, target_local, decomp, is_ntt, first, count, send, stream=None):
        _chk(self.L.b2_keyswitch_shard_prepare(self.h, _dptr(target_local), decomp, int(is_ntt), first, count, _dptr(send), self._stream(stream)))

    def keyswitch_shard_products(self, gathered, world, pad, decomp, first, count, shard, u, stream=None):
        _chk(self.L.b2_keyswitch_shard_products(self.h, _dptr(gathered), world, pad, decomp, first, count, shard.h, _dptr(u), self._stream(stream)))

    def keyswitch_shard_finish(self, u, decomp, is_ntt, first, count, dest_local, assign=ASSIGN_OVERWRITE, stream=None):
        _chk(self.L.b2_keyswitch_shard_finish(self.h, _dptr(u), decomp, int(is_ntt), first, count, _dptr(dest_local), assign, self._stream(stream)))

    # R1-R4 -------------------------------------------------------------------------------------
    def mod_switch_to_next(self, src, dst, count, nl, kind, stream=None):
        _chk(self.L.b2_mod_switch_to_next(self.h, _dptr(src), _dptr(dst), count, nl, kind, self._stream(stream)))
        return dst

    # element-wise ------------------------------------------------------------------------------
    def add(self, a, b, out, count, ncomp, table_base=0, stream=None):
        _chk(self.L.b2_add(self.h, _dptr(a), _dptr(b), _dptr(out), count, ncomp, table_base, self._stream(stream))); return out

    def sub(self, a, b, out, count, ncomp, table_base=0, stream=None):
        _chk(self.L.b2_sub(self.h, _dptr(a), _dptr(b), _dptr(out), count, ncomp, table_base, self._stream(stream))); return out

    def negate(self, a, out, count, ncomp, table_base=0, stream=None):
        _chk(self.L.b2_negate(self.h, _dptr(a), _dptr(out), count, ncomp, table_base, self._stream(stream))); return out

    def dyadic_convolute(self, a, na, b, nb, out, ncomp, batch=1, stream=None):
        _chk(self.L.b2_dyadic_convolute(self.h, _dptr(a), na, _dptr(b), nb, _dptr(out), batch, ncomp, self._stream(stream))); return out

    def dyadic_square(self, a, out, ncomp, batch=1, stream=None):
        _chk(self.L.b2_dyadic_square(self.h, _dptr(a), _dptr(out), batch, ncomp, self._stream(stream))); return out

    def dyadic_broadcast(self, a, plain, out, pcount, ncomp, accumulate=False, stream=None):
        _chk(self.L.b2_dyadic_broadcast(self.h, _dptr(a), _dptr(plain), _dptr(out), pcount, ncomp, int(accumulate), self._stream(stream))); return out

    def dyadic_broadcast_batched(self, a, plain, out, batch, pcount, ncomp, accumulate=False, stream=None):
        _chk(self.L.b2_dyadic_broadcast_batched(self.h, _dptr(a), _dptr(plain), _dptr(out), batch, pcount, ncomp, int(accumulate), self._stream(stream))); return out

    def multiply_plain_accumulate(self, cts, pts, dests, pcount, ncomp, add_to_dest=False, table_base=0, stream=None):
        """dest[i] (+)= ct[i] o pt[i] over lists of device tensors / pointers; destinations may repeat; one launch"""
        nt = len(cts)
        cp = (vp * nt)(*[_dptr(t) for t in cts]); pp = (vp * nt)(*[_dptr(t) for t in pts]); dp = (vp * nt)(*[_dptr(t) for t in dests])
        _chk(self.L.b2_multiply_plain_accumulate(self.h, cp, pp, dp, nt, pcount, ncomp, table_base, int(add_to_dest), self._stream(stream)))
        return dests

    def bfv_multiply(self, a, b, out, nl, batch=1, stream=None):
        _chk(self.L.b2_bfv_multiply(self.h, _dptr(a), _dptr(b), _dptr(out), batch, nl, self._stream(stream))); return out

    def behz_bases(self, nl):
        bsk = (C.c_uint64 * (self.K + 2))(); cnt = C.c_size_t(); g = C.c_uint64()
        _chk(self.L.b2_behz_bases(self.h, nl, bsk, C.byref(cnt), C.byref(g)))
        return [int(bsk[i]) for i in range(cnt.value)], int(g.value)

    # ---- RLWE packing (src/evaluator_lwes.cu:305-491)
    def pack_layer(self, inp, out, target, pairs, nl, shift, galois_element, key, stream=None):
        _chk(self.L.b2_pack_layer(self.h, _dptr(inp), _dptr(out), _dptr(target), pairs, nl, shift, galois_element, key.h, self._stream(stream))); return out

    def pack_rlwes(self, ciphers, input_ntt, nl, shift, input_interval, output_interval, keys, out, apply_field_trace=True, output_ntt=None, stream=None):
        """one group: ciphers = list of device tensors, out = device tensor.  Several groups packed side by side: ciphers = list of lists,
        out = list of tensors.  keys: {galois element: KSwitchKey}; output_ntt defaults to the reference's rule (CKKS / BGV: NTT form)"""
        output_ntt = (self.scheme != BFV) if output_ntt is None else output_ntt
        grouped = len(ciphers) > 0 and isinstance(ciphers[0], (list, tuple))
        groups = ciphers if grouped else [ciphers]
        outs = out if grouped else [out]
        flat = [c for g in groups for c in g]
        counts = (C.c_size_t * len(groups))(*[len(g) for g in groups])
        elts = sorted(keys)
        ea = (C.c_uint64 * len(elts))(*elts); ka = (vp * len(elts))(*[keys[e].h for e in elts])
        _chk(self.L.b2_pack_rlwes(self.h, self._ptrs(flat), counts, len(groups), int(input_ntt), nl, shift, input_interval, output_interval, ea, ka, len(elts),
                                  int(apply_field_trace), int(output_ntt), self._ptrs(outs), self._stream(stream)))
        return out

    def extract_lwe(self, ct, term, nl, c0, c1, stream=None):
        _chk(self.L.b2_extract_lwe(self.h, _dptr(ct), term, nl, _dptr(c0), _dptr(c1), self._stream(stream))); return c0, c1

    def assemble_lwes(self, c0s, c1s, outs, nl, stream=None):
        _chk(self.L.b2_assemble_lwes(self.h, self._ptrs(c0s), self._ptrs(c1s), self._ptrs(outs), len(c0s), nl, self._stream(stream))); return outs

    def pack_lwes(self, groups, nl, keys, outs, apply_field_trace=True, output_ntt=None, stream=None):
        """groups: list of groups, each a list of (c0, c1) device tensors; outs: one device tensor per group"""
        output_ntt = (self.scheme != BFV) if output_ntt is None else output_ntt
        c0s = [l[0] for g in groups for l in g]; c1s = [l[1] for g in groups for l in g]
        counts = (C.c_size_t * len(groups))(*[len(g) for g in groups])
        elts = sorted(keys)
        ea = (C.c_uint64 * len(elts))(*elts); ka = (vp * len(elts))(*[keys[e].h for e in elts])
        _chk(self.L.b2_pack_lwes(self.h, self._ptrs(c0s), self._ptrs(c1s), counts, len(groups), nl, ea, ka, len(elts), int(apply_field_trace), int(output_ntt),
                                 self._ptrs(outs), self._stream(stream)))
        return outs

    # ---- generator, samplers, Encryptor (src/utils/random_generator.cu, src/utils/rlwe.cu, src/encryptor.cu)
    def prng_fill(self, g, out, nbytes, stream=None):
        _chk(self.L.b2_prng_fill(self.h, C.byref(g), _dptr(out), nbytes, self._stream(stream))); return out

    def prng_sample_uint64(self, g):
        return int(self.L.b2_prng_sample_uint64(C.byref(g)))

    def sample_poly(self, kind, g, out, nl, count=1, stream=None):
        _chk(self.L.b2_sample_poly(self.h, kind, C.byref(g), _dptr(out), count, nl, self._stream(stream))); return out

    def encrypt_zero_asymmetric(self, pk, nl, is_ntt, ct, u_prng, e_prng=None, stream=None):
        e_prng = u_prng if e_prng is None else e_prng
        _chk(self.L.b2_encrypt_zero_asymmetric(self.h, _dptr(pk), nl, int(is_ntt), C.byref(u_prng), C.byref(e_prng), _dptr(ct), self._stream(stream))); return ct

    def encrypt_zero_symmetric(self, sk, nl, is_ntt, ct, c1_prng, e_prng=None, save_seed=False, stream=None):
        e_prng = c1_prng if e_prng is None else e_prng
        seed = C.c_uint64(0)
        _chk(self.L.b2_encrypt_zero_symmetric(self.h, _dptr(sk), nl, int(is_ntt), int(save_seed), C.byref(c1_prng), C.byref(e_prng), _dptr(ct), C.byref(seed), self._stream(stream)))
        return int(seed.value)

    def batch_encode(self, values, count, plain, stream=None):
        _chk(self.L.b2_batch_encode(self.h, _dptr(values), count, _dptr(plain), self._stream(stream))); return plain

    def batch_decode(self, plain, plain_count, values, stream=None):
        _chk(self.L.b2_batch_decode(self.h, _dptr(plain), plain_count, _dptr(values), self._stream(stream))); return values

    def generate_secret_key(self, g, sk, stream=None):
        _chk(self.L.b2_generate_secret_key(self.h, C.byref(g), _dptr(sk), self._stream(stream))); return sk

    def generate_kswitch_key(self, sk, new_key, g, out, stream=None):
        """out [K-1][2][K][N]: KeyGenerator::generate_one_kswitch_key for new_key (s^2, or the Galois image of s)"""
        _chk(self.L.b2_generate_kswitch_key(self.h, _dptr(sk), _dptr(new_key), C.byref(g), _dptr(out), self._stream(stream))); return out

    def bfv_scale_up(self, plain, plain_count, nl, dest, frm=None, subtract=False, stream=None):
        _chk(self.L.b2_bfv_scale_up(self.h, _dptr(plain), plain_count, nl, None if frm is None else _dptr(frm), int(subtract), _dptr(dest), self._stream(stream))); return dest

    def centralize(self, plain, plain_count, nl, out, scalar=1, stream=None):
        if scalar == 1:
            _chk(self.L.b2_centralize(self.h, _dptr(plain), plain_count, nl, _dptr(out), self._stream(stream)))
        else:
            _chk(self.L.b2_centralize_scaled(self.h, _dptr(plain), plain_count, nl, scalar, _dptr(out), self._stream(stream)))
        return out

    def encrypt(self, plain, plain_count, key, asymmetric, nl, ct, u_prng, e_prng=None, stream=None):
        e_prng = u_prng if e_prng is None else e_prng
        _chk(self.L.b2_encrypt(self.h, _dptr(plain), plain_count, _dptr(key), int(asymmetric), nl, C.byref(u_prng), C.byref(e_prng), _dptr(ct), self._stream(stream))); return ct

    # ---- Decryptor (src/decryptor.cu): lists of device tensors / pointers
    def secret_key_powers(self, sk, max_power, out, stream=None):
        _chk(self.L.b2_secret_key_powers(self.h, _dptr(sk), max_power, _dptr(out), self._stream(stream))); return out

    @staticmethod
    def _ptrs(ts):
        return (vp * len(ts))(*[_dptr(t) for t in ts])

    def decrypt_phase(self, cts, outs, pcount, nl, is_ntt, sk_powers, stream=None):
        _chk(self.L.b2_decrypt_phase(self.h, self._ptrs(cts), self._ptrs(outs), len(cts), pcount, nl, int(is_ntt), _dptr(sk_powers), self._stream(stream))); return outs

    def decrypt_scale_and_round(self, phases, outs, nl, stream=None):
        _chk(self.L.b2_decrypt_scale_and_round(self.h, self._ptrs(phases), self._ptrs(outs), len(phases), nl, self._stream(stream))); return outs

    def decrypt_mod_t(self, phases, outs, nl, correction_factors=None, stream=None):
        cf = None if correction_factors is None else (C.c_uint64 * len(phases))(*[int(c) for c in correction_factors])
        _chk(self.L.b2_decrypt_mod_t(self.h, self._ptrs(phases), self._ptrs(outs), len(phases), nl, cf, self._stream(stream))); return outs

    def decrypt(self, cts, outs, pcount, nl, sk_powers, correction_factors=None, stream=None):
        """Decryptor::decrypt_batched: BFV / BGV -> [N] words mod t per ciphertext, CKKS -> [nl][N] (NTT form)"""
        cf = None if correction_factors is None else (C.c_uint64 * len(cts))(*[int(c) for c in correction_factors])
        _chk(self.L.b2_decrypt(self.h, self._ptrs(cts), self._ptrs(outs), len(cts), pcount, nl, _dptr(sk_powers), cf, self._stream(stream))); return outs

    def scale_inv_degree(self, data, count, ncomp, scalar=1, table_base=0, stream=None):
        _chk(self.L.b2_scale_inv_degree(self.h, _dptr(data), count, ncomp, table_base, scalar, self._stream(stream))); return data

    def negacyclic_shift(self, a, shift, out, count, ncomp, table_base=0, stream=None):
        _chk(self.L.b2_negacyclic_shift(self.h, _dptr(a), shift, _dptr(out), count, ncomp, table_base, self._stream(stream))); return out

    def galois(self, src, galois_elt, out, count, ncomp, is_ntt, table_base=0, stream=None):
        _chk(self.L.b2_galois(self.h, _dptr(src), galois_elt, _dptr(out), count, ncomp, table_base, int(is_ntt), self._stream(stream))); return out


class KSwitchKey:
    def __init__(self, ctx, h, keep):
        self.ctx, self.h, self.keep = ctx, h, keep

    def __del__(self):
        try:
            if self.h:
                lib().b2_kswitch_key_destroy(self.h); self.h = None
        except Exception:
            pass
