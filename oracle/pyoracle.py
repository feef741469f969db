"""ctypes front-ends for the CPU oracle (oracle/_build/librns_oracle.so) and, when built, the
unmodified reference behind oracle/_ref/libref_shim.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never by the product package (troy-nova_b200/).
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "librns_oracle.so")
REF_SHIM_SO = os.path.join(HERE, "_ref", "libref_shim.so")

BFV, CKKS, BGV = 1, 2, 3
ASSIGN_OVERWRITE, ASSIGN_ADD_INPLACE, ASSIGN_OVERWRITE_EXCEPT_FIRST = 0, 1, 2
TABLES_COMPONENTWISE, TABLES_KS_SET_PRODUCTS, TABLES_KS_SKIP_FINALS = 0, 1, 2

u64p = C.POINTER(C.c_uint64)


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def build_oracle():
    """Compile the C restatement (gcc, seconds)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def build_ref():
    """Compile the unmodified reference + shim; needs /root/reference (this container only)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.ro_ctx_create.restype = C.c_void_p
        L.ro_ctx_create.argtypes = [C.c_uint, u64p, C.c_size_t, C.c_uint64, C.c_int]
        L.ro_ctx_destroy.argtypes = [C.c_void_p]
        for f in ("ro_ctx_modulus", "ro_ctx_root"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p, C.c_size_t]
        for f in ("ro_ctx_root_powers", "ro_ctx_inv_root_powers"):
            getattr(L, f).restype = u64p
            getattr(L, f).argtypes = [C.c_void_p, C.c_size_t]
        L.ro_min_primitive_root.restype = C.c_uint64
        L.ro_min_primitive_root.argtypes = [C.c_uint64, C.c_uint64]
        L.ro_plain_modulus_batching.restype = C.c_uint64
        L.ro_plain_modulus_batching.argtypes = [C.c_size_t, C.c_uint]
        L.ro_coeff_modulus_create.argtypes = [C.c_size_t, C.POINTER(C.c_uint), C.c_size_t, u64p]
        L.ro_get_primes.argtypes = [C.c_uint64, C.c_uint, C.c_size_t, u64p]
        for f in ("ro_mulmod", "ro_powmod"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64] * 3
        L.ro_invmod.restype = C.c_uint64
        L.ro_invmod.argtypes = [C.c_uint64] * 2
        L.ro_barrett_reduce64.restype = C.c_uint64
        L.ro_barrett_reduce64.argtypes = [C.c_uint64] * 2
        L.ro_barrett_reduce128.restype = C.c_uint64
        L.ro_barrett_reduce128.argtypes = [C.c_uint64] * 3
        L.ro_shoup_quotient.restype = C.c_uint64
        L.ro_shoup_quotient.argtypes = [C.c_uint64] * 2
        L.ro_shoup_mul_lazy.restype = C.c_uint64
        L.ro_shoup_mul_lazy.argtypes = [C.c_uint64] * 4
        L.ro_is_prime.argtypes = [C.c_uint64]
        ntt_args = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t]
        L.ro_ntt_fwd.argtypes = ntt_args
        L.ro_ntt_inv.argtypes = ntt_args
        L.ro_add.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t, C.c_size_t]
        L.ro_sub.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t, C.c_size_t]
        L.ro_negate.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_size_t]
        L.ro_dyadic_convolute.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t, u64p, C.c_size_t]
        L.ro_dyadic_square.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        L.ro_dyadic_broadcast.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t, C.c_size_t, C.c_int]
        L.ro_negacyclic_shift.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t, C.c_size_t]
        L.ro_galois_element_from_step.restype = C.c_uint64
        L.ro_galois_element_from_step.argtypes = [C.c_size_t, C.c_int]
        L.ro_galois_ntt_table.argtypes = [C.c_uint, C.c_uint64, C.POINTER(C.c_uint32)]
        L.ro_galois_coeff.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, C.c_size_t, C.c_size_t]
        L.ro_galois_ntt.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, C.c_size_t, C.c_size_t]
        L.ro_keyswitch.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int, C.POINTER(u64p), u64p, C.c_int]
        for f in ("ro_divide_round_q_last", "ro_divide_round_q_last_ntt", "ro_mod_t_divide_q_last_ntt", "ro_drop_last_limb"):
            getattr(L, f).argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_size_t]
        L.ro_ntt_multiply_inv_degree.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64]
        L.ro_bfv_multiply.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t]
        L.ro_behz_bases.argtypes = [C.c_void_p, C.c_size_t, u64p, C.POINTER(C.c_size_t), u64p]
        L.ro_keygen_secret.argtypes = [C.c_void_p, C.c_uint64, u64p]
        L.ro_keygen_kswitch.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p]
        L.ro_encrypt_zero_sym.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int, C.c_uint64, u64p]
        L.ro_phase.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_size_t, C.c_int, u64p]
        L.ro_secret_key_powers.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.ro_decrypt_phase.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_int, u64p, u64p]
        L.ro_decrypt_scale_and_round.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.ro_decrypt_mod_t.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_uint64, u64p]
        L.ro_decrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p, C.c_uint64, u64p]
        u8p = C.POINTER(C.c_uint8)
        L.ro_prng_fill_bytes.argtypes = [u64p, u8p, C.c_size_t]
        L.ro_prng_sample_uint64.restype = C.c_uint64
        L.ro_prng_sample_uint64.argtypes = [u64p]
        for f in ("ro_sample_poly_ternary", "ro_sample_poly_cbd", "ro_sample_poly_uniform"):
            getattr(L, f).argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        L.ro_encrypt_zero_asymmetric.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int, u64p, u64p, u64p]
        L.ro_encrypt_zero_symmetric.restype = C.c_uint64
        L.ro_encrypt_zero_symmetric.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int, C.c_int, u64p, u64p, u64p]
        L.ro_bfv_scale_up.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p, C.c_int, u64p]
        L.ro_centralize.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p]
        L.ro_encrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_int, C.c_size_t, u64p, u64p, u64p]
        L.ro_keygen_secret_prng.argtypes = [C.c_void_p, u64p, u64p]
        L.ro_keygen_kswitch_prng.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p]
        L.ro_fnv1a64.restype = C.c_uint64
        L.ro_fnv1a64.argtypes = [u64p, C.c_size_t]
        _lib = L
    return _lib


def coeff_modulus_create(n, bits):
    out = np.zeros(len(bits), dtype=np.uint64)
    arr = (C.c_uint * len(bits))(*bits)
    if lib().ro_coeff_modulus_create(n, arr, len(bits), _p(out)):
        raise ValueError("ro_coeff_modulus_create failed")
    return out


def plain_modulus_batching(n, bits):
    return int(lib().ro_plain_modulus_batching(n, bits))


def fnv1a64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
    return int(lib().ro_fnv1a64(_p(a), a.size))


class Oracle:
    """CPU oracle context: key-level moduli (special prime last), plain modulus, scheme."""

    def __init__(self, log_n, moduli, plain_modulus=0, scheme=CKKS):
        self.L = lib()
        self.log_n, self.n = log_n, 1 << log_n
        self.moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
        self.nm = len(self.moduli)
        self.t, self.scheme = int(plain_modulus), scheme
        self.h = self.L.ro_ctx_create(log_n, _p(self.moduli), self.nm, self.t, scheme)
        if not self.h:
            raise ValueError("ro_ctx_create failed (moduli must be primes = 1 mod 2N)")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ro_ctx_destroy(self.h)
            self.h = None

    def root(self, i):
        return int(self.L.ro_ctx_root(self.h, i))

    def root_powers(self, i, inverse=False):
        f = self.L.ro_ctx_inv_root_powers if inverse else self.L.ro_ctx_root_powers
        return np.ctypeslib.as_array(f(self.h, i), shape=(self.n, 2)).copy()

    def _ntt(self, fn, data, pcount, ncomp, mode, base, decomp):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        assert d.size == pcount * ncomp * self.n
        fn(self.h, _p(d), pcount, ncomp, mode, base, decomp)
        return d

    def ntt_fwd(self, data, pcount, ncomp, mode=TABLES_COMPONENTWISE, base=0, decomp=0):
        return self._ntt(self.L.ro_ntt_fwd, data, pcount, ncomp, mode, base, decomp)

    def ntt_inv(self, data, pcount, ncomp, mode=TABLES_COMPONENTWISE, base=0, decomp=0):
        return self._ntt(self.L.ro_ntt_inv, data, pcount, ncomp, mode, base, decomp)

    def add(self, a, b, pc, nc):
        o = np.empty_like(a); self.L.ro_add(self.h, _p(a), _p(b), _p(o), pc, nc); return o

    def sub(self, a, b, pc, nc):
        o = np.empty_like(a); self.L.ro_sub(self.h, _p(a), _p(b), _p(o), pc, nc); return o

    def negate(self, a, pc, nc):
        o = np.empty_like(a); self.L.ro_negate(self.h, _p(a), _p(o), pc, nc); return o

    def dyadic_convolute(self, a, na, b, nb, nc):
        o = np.zeros((na + nb - 1) * nc * self.n, dtype=np.uint64)
        self.L.ro_dyadic_convolute(self.h, _p(a), na, _p(b), nb, _p(o), nc); return o

    def dyadic_square(self, a, nc):
        o = np.zeros(3 * nc * self.n, dtype=np.uint64)
        self.L.ro_dyadic_square(self.h, _p(a), _p(o), nc); return o

    def dyadic_broadcast(self, a, plain, pc, nc, acc_into=None):
        o = np.zeros(pc * nc * self.n, dtype=np.uint64) if acc_into is None else acc_into.copy()
        self.L.ro_dyadic_broadcast(self.h, _p(a), _p(plain), _p(o), pc, nc, 0 if acc_into is None else 1); return o

    def negacyclic_shift(self, a, shift, pc, nc):
        o = np.empty_like(a); self.L.ro_negacyclic_shift(self.h, _p(a), shift, _p(o), pc, nc); return o

    def galois_element_from_step(self, step):
        return int(self.L.ro_galois_element_from_step(self.n, step))

    def galois_ntt_table(self, g):
        t = np.zeros(self.n, dtype=np.uint32)
        self.L.ro_galois_ntt_table(self.log_n, g, t.ctypes.data_as(C.POINTER(C.c_uint32))); return t

    def galois(self, a, g, pc, nc, ntt_form):
        o = np.empty_like(a)
        (self.L.ro_galois_ntt if ntt_form else self.L.ro_galois_coeff)(self.h, _p(a), g, _p(o), pc, nc); return o

    def keyswitch(self, target, decomp, is_ntt, key, dest=None, assign=ASSIGN_OVERWRITE):
        """key: array [decomp_first][2][K][N]; dest [2][decomp][N] (copied) -> result."""
        key = np.ascontiguousarray(key, dtype=np.uint64).reshape(-1)
        per = 2 * self.nm * self.n
        ptrs = (u64p * decomp)(*[C.cast(key.ctypes.data + j * per * 8, u64p) for j in range(decomp)])
        d = np.zeros(2 * decomp * self.n, dtype=np.uint64) if dest is None else np.ascontiguousarray(dest, dtype=np.uint64).copy()
        t = np.ascontiguousarray(target, dtype=np.uint64)
        self.L.ro_keyswitch(self.h, _p(t), decomp, int(is_ntt), ptrs, _p(d), assign)
        return d

    def _modswitch(self, fn, a, pc, nl):
        o = np.zeros(pc * (nl - 1) * self.n, dtype=np.uint64)
        fn(self.h, _p(np.ascontiguousarray(a, dtype=np.uint64)), _p(o), pc, nl); return o

    def divide_round_q_last(self, a, pc, nl): return self._modswitch(self.L.ro_divide_round_q_last, a, pc, nl)
    def divide_round_q_last_ntt(self, a, pc, nl): return self._modswitch(self.L.ro_divide_round_q_last_ntt, a, pc, nl)
    def mod_t_divide_q_last_ntt(self, a, pc, nl): return self._modswitch(self.L.ro_mod_t_divide_q_last_ntt, a, pc, nl)
    def drop_last_limb(self, a, pc, nl): return self._modswitch(self.L.ro_drop_last_limb, a, pc, nl)

    def bfv_multiply(self, a, b, nl):
        o = np.zeros(3 * nl * self.n, dtype=np.uint64)
        if self.L.ro_bfv_multiply(self.h, _p(a), _p(b), _p(o), nl):
            raise NotImplementedError("ro_bfv_multiply")
        return o

    def behz_bases(self, nl):
        bsk = np.zeros(self.nm + 2, dtype=np.uint64); cnt = C.c_size_t(); g = np.zeros(1, dtype=np.uint64)
        if self.L.ro_behz_bases(self.h, nl, _p(bsk), C.byref(cnt), _p(g)):
            raise ValueError("ro_behz_bases")
        return [int(v) for v in bsk[:cnt.value]], int(g[0])

    def ntt_multiply_inv_degree(self, data, pc, nc, scalar=1, base=0):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        self.L.ro_ntt_multiply_inv_degree(self.h, _p(d), pc, nc, base, scalar); return d

    # toy RLWE -----------------------------------------------------------
    def keygen_secret(self, seed):
        sk = np.zeros(self.nm * self.n, dtype=np.uint64); self.L.ro_keygen_secret(self.h, seed, _p(sk)); return sk

    def keygen_kswitch(self, sk, new_key, seed):
        o = np.zeros((self.nm - 1) * 2 * self.nm * self.n, dtype=np.uint64)
        self.L.ro_keygen_kswitch(self.h, _p(sk), _p(new_key), seed, _p(o)); return o

    def encrypt_zero_sym(self, sk, nl, is_ntt, seed):
        ct = np.zeros(2 * nl * self.n, dtype=np.uint64)
        self.L.ro_encrypt_zero_sym(self.h, _p(sk), nl, int(is_ntt), seed, _p(ct)); return ct

    def phase(self, sk, ct, pc, nl, is_ntt):
        o = np.zeros(nl * self.n, dtype=np.uint64)
        self.L.ro_phase(self.h, _p(sk), _p(np.ascontiguousarray(ct, dtype=np.uint64)), pc, nl, int(is_ntt), _p(o)); return o

    # ---- Decryptor
    def secret_key_powers(self, sk, max_power):
        o = np.zeros(max_power * self.nm * self.n, dtype=np.uint64)
        self.L.ro_secret_key_powers(self.h, _p(np.ascontiguousarray(sk, dtype=np.uint64)), max_power, _p(o)); return o

    def decrypt_phase(self, ct, pc, nl, is_ntt, sk_powers):
        o = np.zeros(nl * self.n, dtype=np.uint64)
        self.L.ro_decrypt_phase(self.h, _p(np.ascontiguousarray(ct, dtype=np.uint64)), pc, nl, int(is_ntt), _p(sk_powers), _p(o)); return o

    def decrypt_scale_and_round(self, phase, nl):
        o = np.zeros(self.n, dtype=np.uint64)
        if self.L.ro_decrypt_scale_and_round(self.h, _p(np.ascontiguousarray(phase, dtype=np.uint64)), nl, _p(o)): raise RuntimeError("ro_decrypt_scale_and_round failed")
        return o

    def decrypt_mod_t(self, phase, nl, correction_factor=1):
        o = np.zeros(self.n, dtype=np.uint64)
        if self.L.ro_decrypt_mod_t(self.h, _p(np.ascontiguousarray(phase, dtype=np.uint64)), nl, correction_factor, _p(o)): raise RuntimeError("ro_decrypt_mod_t failed")
        return o

    def decrypt(self, ct, pc, nl, sk_powers, correction_factor=1):
        o = np.zeros((nl if self.scheme == CKKS else 1) * self.n, dtype=np.uint64)
        if self.L.ro_decrypt(self.h, _p(np.ascontiguousarray(ct, dtype=np.uint64)), pc, nl, _p(sk_powers), correction_factor, _p(o)): raise RuntimeError("ro_decrypt failed")
        return o

    # ---- RLWE packing: Evaluator::pack_rlwe_ciphertexts_new, device branch, step by step (src/evaluator_lwes.cu:305-491, field trace :96-105)
    def apply_galois(self, ct, g, nl, is_ntt, key):
        """Evaluator::apply_galois: permute both polynomials, keyswitch c_1 with OverwriteExceptFirst (src/evaluator_keyswitching.cu:143-200)"""
        ga = self.galois(ct, g, 2, nl, is_ntt)
        return self.keyswitch(ga[nl * self.n:], nl, is_ntt, key, dest=ga, assign=ASSIGN_OVERWRITE_EXCEPT_FIRST)

    def pack_rlwes(self, ciphers, input_ntt, nl, shift, input_interval, output_interval, keys, apply_field_trace=True):
        """ciphers: list of [2][nl][N] arrays; keys: {galois element: key array}.  Output NTT form for CKKS / BGV, coefficient form for BFV."""
        n = self.n
        out_ntt = self.scheme != BFV
        M = input_interval // output_interval
        layers = M.bit_length() - 1
        assert 1 <= len(ciphers) <= M
        w = 2 * nl * n
        rl = []
        for i in range(M):
            idx = int(format(i, "0%db" % layers)[::-1], 2) if layers else 0
            if idx < len(ciphers):
                c = np.ascontiguousarray(ciphers[idx], dtype=np.uint64).copy()
                if input_ntt:
                    c = self.ntt_inv(c, 2, nl)
                c = self.ntt_multiply_inv_degree(c, 2, nl, scalar=n // input_interval)
                if shift:
                    c = self.negacyclic_shift(c, shift, 2, nl)
                rl.append(c)
            else:
                rl.append(np.zeros(w, dtype=np.uint64))
        for layer in range(layers):
            gap = 1 << layer
            sh = input_interval >> (layer + 1)
            g = (n // input_interval) * (1 << (layer + 1)) + 1
            for off in range(0, M, 2 * gap):
                even, odd = rl[off], rl[off + gap]
                temp = self.negacyclic_shift(odd, sh, 2, nl)
                odd = self.sub(even, temp, 2, nl)
                even = self.add(even, temp, 2, nl)
                if out_ntt:
                    odd = self.ntt_fwd(odd, 2, nl)
                odd = self.apply_galois(odd, g, nl, out_ntt, keys[g])
                if out_ntt:
                    odd = self.ntt_inv(odd, 2, nl)
                rl[off] = self.add(even, odd, 2, nl)
        ret = rl[0]
        if out_ntt:
            ret = self.ntt_fwd(ret, 2, nl)
        if output_interval != 1 and apply_field_trace:
            deg = n
            while deg > n // output_interval:
                ret = self.add(ret, self.apply_galois(ret, deg + 1, nl, out_ntt, keys[deg + 1]), 2, nl)
                deg >>= 1
        return ret

    # ---- BatchEncoder (src/batch_encoder.cu:14-60,169-227,407-460): slot i <-> coefficient index_map[i] of the NTT-form polynomial modulo t
    def _batching(self):
        if not hasattr(self, "_bt"):
            n, logn = self.n, self.log_n
            row, m, pos = n >> 1, n << 1, 1
            mp = np.zeros(n, dtype=np.int64)
            rev = lambda v: int(format(v, "0%db" % logn)[::-1], 2)
            for i in range(row):
                mp[i] = rev((pos - 1) >> 1); mp[i + row] = rev((m - pos - 1) >> 1)
                pos = (pos * 3) & (m - 1)
            self._bt = (mp, Oracle(logn, np.array([self.t], dtype=np.uint64), 0, BFV))      # negacyclic NTT tables modulo t (minimal primitive root, as plain_ntt_tables)
        return self._bt

    def batch_encode(self, values):
        mp, ot = self._batching()
        a = np.zeros(self.n, dtype=np.uint64)
        v = np.ascontiguousarray(values, dtype=np.uint64)
        a[mp[:v.size]] = v
        return ot.ntt_inv(a, 1, 1)

    def batch_decode(self, plain):
        mp, ot = self._batching()
        a = np.zeros(self.n, dtype=np.uint64)
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        a[:p.size] = p
        return ot.ntt_fwd(a, 1, 1)[mp]

    # ---- LWE (src/evaluator_lwes.cu:51-94, src/lwe_ciphertext.cu:12-140, pack_lwe_ciphertexts_new :201-229)
    def extract_lwe(self, ct, term, nl):
        """coefficient-form ct [2][nl][N] -> (c0 [nl], c1 [nl][N])"""
        n = self.n
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        c1 = self.negacyclic_shift(ct[nl * n:], 0 if term == 0 else 2 * n - term, 1, nl)
        return ct[:nl * n].reshape(nl, n)[:, term].copy(), c1

    def assemble_lwe(self, c0, c1, nl):
        n = self.n
        out = np.zeros(2 * nl * n, dtype=np.uint64)
        out[:nl * n].reshape(nl, n)[:, 0] = c0
        out[nl * n:] = c1
        return out

    def pack_lwes(self, lwes, nl, keys, apply_field_trace=True):
        l = 0
        while (1 << l) < len(lwes):
            l += 1
        rl = [self.assemble_lwe(c0, c1, nl) for c0, c1 in lwes]
        return self.pack_rlwes(rl, False, nl, 0, self.n, self.n >> l, keys, apply_field_trace)

    # ---- generator (state = np.uint64[4]: seed_lo, seed_hi, counter_lo, counter_hi; advanced in place), samplers, Encryptor
    @staticmethod
    def prng_state(seed_lo, seed_hi=0, counter=0):
        return np.array([seed_lo, seed_hi, counter, 0], dtype=np.uint64)

    def prng_fill_bytes(self, st, nbytes):
        o = np.zeros(nbytes, dtype=np.uint8)
        self.L.ro_prng_fill_bytes(_p(st), o.ctypes.data_as(C.POINTER(C.c_uint8)), nbytes); return o

    def sample_poly(self, kind, st, nl):
        o = np.zeros(nl * self.n, dtype=np.uint64)
        {"ternary": self.L.ro_sample_poly_ternary, "cbd": self.L.ro_sample_poly_cbd, "uniform": self.L.ro_sample_poly_uniform}[kind](self.h, _p(st), _p(o), nl); return o

    def encrypt_zero_asymmetric(self, pk, nl, is_ntt, u_st, e_st=None):
        ct = np.zeros(2 * nl * self.n, dtype=np.uint64)
        e_st = u_st if e_st is None else e_st
        if self.L.ro_encrypt_zero_asymmetric(self.h, _p(pk), nl, int(is_ntt), _p(u_st), _p(e_st), _p(ct)): raise RuntimeError("ro_encrypt_zero_asymmetric failed")
        return ct

    def encrypt_zero_symmetric(self, sk, nl, is_ntt, save_seed, c1_st, e_st=None):
        ct = np.zeros(2 * nl * self.n, dtype=np.uint64)
        e_st = c1_st if e_st is None else e_st
        seed = self.L.ro_encrypt_zero_symmetric(self.h, _p(sk), nl, int(is_ntt), int(save_seed), _p(c1_st), _p(e_st), _p(ct))
        return ct, int(seed)

    def keygen_secret_prng(self, st):
        sk = np.zeros(self.nm * self.n, dtype=np.uint64)
        self.L.ro_keygen_secret_prng(self.h, _p(st), _p(sk)); return sk

    def keygen_kswitch_prng(self, sk, new_key, st):
        out = np.zeros(max(self.nm - 1, 1) * 2 * self.nm * self.n, dtype=np.uint64)
        self.L.ro_keygen_kswitch_prng(self.h, _p(np.ascontiguousarray(sk, dtype=np.uint64)), _p(np.ascontiguousarray(new_key, dtype=np.uint64)), _p(st), _p(out)); return out

    def bfv_scale_up(self, plain, nl, frm=None, subtract=False):
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        o = np.zeros(nl * self.n, dtype=np.uint64)
        self.L.ro_bfv_scale_up(self.h, _p(plain), plain.size, nl, None if frm is None else _p(np.ascontiguousarray(frm, dtype=np.uint64)), int(subtract), _p(o)); return o

    def centralize(self, plain, nl):
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        o = np.zeros(nl * self.n, dtype=np.uint64)
        self.L.ro_centralize(self.h, _p(plain), plain.size, nl, _p(o)); return o

    def plain_op(self, op, ct, nl, plain, correction_factor=1):
        """Evaluator::multiply_plain (op 0) / add_plain (1) / sub_plain (2) with a plaintext modulo t on a ciphertext in the scheme's default form
        (src/evaluator_multiply_plain.cu:13-68, src/evaluator_translate_plain.cu:95-170)"""
        n, w = self.n, nl * self.n
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        if op == 0:
            m = self.ntt_fwd(self.centralize(plain, nl), 1, nl)
            if self.scheme == BFV:
                return self.ntt_inv(self.dyadic_broadcast(self.ntt_fwd(ct, 2, nl), m, 2, nl), 2, nl)
            return self.dyadic_broadcast(ct, m, 2, nl)
        f = self.sub if op == 2 else self.add
        if self.scheme == BFV:
            return np.concatenate([self.bfv_scale_up(plain, nl, ct[:w], op == 2), ct[w:]])
        scaled = plain if correction_factor == 1 else np.array([(int(v) * correction_factor) % self.t for v in plain], dtype=np.uint64)
        m = self.ntt_fwd(self.centralize(scaled, nl), 1, nl)
        return np.concatenate([f(ct[:w], m, 1, nl), ct[w:]])

    def encrypt(self, plain, key, asymmetric, nl, u_st, e_st=None):
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        ct = np.zeros(2 * nl * self.n, dtype=np.uint64)
        e_st = u_st if e_st is None else e_st
        cnt = plain.size if self.scheme != CKKS else self.n
        if self.L.ro_encrypt(self.h, _p(plain), cnt, _p(key), int(asymmetric), nl, _p(u_st), _p(e_st), _p(ct)): raise RuntimeError("ro_encrypt failed")
        return ct

    def random_poly(self, rng, pc, nl, base=0):
        """uniform residues, [pc][nl][N], limb j reduced mod moduli[base+j]"""
        out = np.empty((pc, nl, self.n), dtype=np.uint64)
        for j in range(nl):
            out[:, j, :] = rng.integers(0, int(self.moduli[base + j]), size=(pc, self.n), dtype=np.uint64)
        return out.reshape(-1)

    def random_key(self, rng, decomp=None):
        """uniform 'key' [decomp][2][K][N] (structure-free; keyswitch is a deterministic function of it)"""
        decomp = self.nm - 1 if decomp is None else decomp
        out = np.empty((decomp, 2, self.nm, self.n), dtype=np.uint64)
        for i in range(self.nm):
            out[:, :, i, :] = rng.integers(0, int(self.moduli[i]), size=(decomp, 2, self.n), dtype=np.uint64)
        return out.reshape(-1)


# ---------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_SHIM_SO)


def reflib():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_SHIM_SO)
        R.rs_create.restype = C.c_void_p
        R.rs_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_uint), C.c_size_t, C.c_uint, C.c_uint64, C.c_int]
        R.rs_error.restype = C.c_char_p
        R.rs_error.argtypes = [C.c_void_p]
        R.rs_destroy.argtypes = [C.c_void_p]
        R.rs_n_key_moduli.restype = C.c_size_t
        R.rs_n_key_moduli.argtypes = [C.c_void_p]
        R.rs_key_moduli.argtypes = [C.c_void_p, u64p]
        R.rs_plain_modulus.restype = C.c_uint64
        R.rs_plain_modulus.argtypes = [C.c_void_p]
        R.rs_ntt_table.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p, C.c_size_t]
        R.rs_gen_relin.argtypes = [C.c_void_p]
        R.rs_gen_galois.argtypes = [C.c_void_p, u64p, C.c_size_t]
        R.rs_relin_key.argtypes = [C.c_void_p, u64p]
        R.rs_galois_key.argtypes = [C.c_void_p, C.c_uint64, u64p]
        R.rs_secret_key.argtypes = [C.c_void_p, u64p]
        R.rs_galois_element_from_step.restype = C.c_uint64
        R.rs_galois_element_from_step.argtypes = [C.c_void_p, C.c_int]
        R.rs_encrypt_zero.argtypes = [C.c_void_p, u64p]
        ntt_args = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t]
        R.rs_ntt_fwd.argtypes = ntt_args
        R.rs_ntt_inv.argtypes = ntt_args
        R.rs_multiply.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t]
        R.rs_square.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        R.rs_relinearize.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        R.rs_apply_galois.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, C.c_size_t]
        R.rs_rotate.argtypes = [C.c_void_p, u64p, C.c_int, u64p, C.c_size_t]
        for f in ("rs_mod_switch_scale_to_next", "rs_drop_to_next", "rs_transform_to_ntt", "rs_transform_from_ntt"):
            getattr(R, f).argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t]
        R.rs_multiply_plain_raw.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, u64p, C.c_size_t, C.c_int]
        R.rs_galois_raw.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_uint64, u64p, C.c_int]
        R.rs_decrypt_raw.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_uint64, u64p]
        R.rs_decrypt_phase_raw.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p]
        R.rs_prng.argtypes = [C.c_void_p, C.c_int, u64p, C.c_size_t, C.c_size_t, u64p]
        R.rs_public_key.argtypes = [C.c_void_p, u64p]
        R.rs_encrypt_seeded.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_uint64, C.c_uint64, u64p, u64p]
        R.rs_pack_rlwes.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, u64p]
        R.rs_extract_lwe.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_int, u64p, u64p]
        R.rs_pack_lwes.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_size_t, C.c_int, u64p]
        R.rs_plain_op.argtypes = [C.c_void_p, C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, u64p]
        R.rs_batch_encode.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        R.rs_batch_decode.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        R.rs_time_op.restype = C.c_double
        R.rs_time_op.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
        _ref = R
    return _ref


class Reference:
    """The unmodified reference (host branch unless device=True) behind the C shim."""

    def __init__(self, scheme, n, bits, t_bits=20, seed=0x123, device=False):
        self.R = reflib()
        self.scheme, self.n = scheme, n
        arr = (C.c_uint * len(bits))(*bits)
        self.h = self.R.rs_create(scheme, n, arr, len(bits), t_bits, seed, int(device))
        err = self.R.rs_error(self.h)
        if err:
            raise RuntimeError(err.decode())
        self.nm = self.R.rs_n_key_moduli(self.h)
        self.moduli = np.zeros(self.nm, dtype=np.uint64)
        self.R.rs_key_moduli(self.h, _p(self.moduli))
        self.t = int(self.R.rs_plain_modulus(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.R.rs_destroy(self.h); self.h = None

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.R.rs_error(self.h).decode())

    def ntt_table(self, i, count):
        root = C.c_uint64(0)
        rp = np.zeros((count, 2), dtype=np.uint64); irp = np.zeros((count, 2), dtype=np.uint64)
        self._chk(self.R.rs_ntt_table(self.h, i, C.byref(root), _p(rp), _p(irp), count))
        return int(root.value), rp, irp

    def gen_relin(self): self._chk(self.R.rs_gen_relin(self.h))

    def gen_galois(self, elts):
        e = np.ascontiguousarray(elts, dtype=np.uint64); self._chk(self.R.rs_gen_galois(self.h, _p(e), e.size))

    def relin_key(self):
        o = np.zeros((self.nm - 1) * 2 * self.nm * self.n, dtype=np.uint64); self._chk(self.R.rs_relin_key(self.h, _p(o))); return o

    def galois_key(self, elt):
        o = np.zeros((self.nm - 1) * 2 * self.nm * self.n, dtype=np.uint64); self._chk(self.R.rs_galois_key(self.h, elt, _p(o))); return o

    def secret_key(self):
        o = np.zeros(self.nm * self.n, dtype=np.uint64); self._chk(self.R.rs_secret_key(self.h, _p(o))); return o

    def galois_element_from_step(self, step): return int(self.R.rs_galois_element_from_step(self.h, step))

    def encrypt_zero(self):
        o = np.zeros(2 * (self.nm - 1) * self.n, dtype=np.uint64); self._chk(self.R.rs_encrypt_zero(self.h, _p(o))); return o

    def ntt_fwd(self, data, pc, nc, mode=0, base=0, decomp=0):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy(); self._chk(self.R.rs_ntt_fwd(self.h, _p(d), pc, nc, mode, base, decomp)); return d

    def ntt_inv(self, data, pc, nc, mode=0, base=0, decomp=0):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy(); self._chk(self.R.rs_ntt_inv(self.h, _p(d), pc, nc, mode, base, decomp)); return d

    def multiply(self, a, b, nl):
        o = np.zeros(3 * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_multiply(self.h, _p(a), _p(b), _p(o), nl)); return o

    def square(self, a, nl):
        o = np.zeros(3 * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_square(self.h, _p(a), _p(o), nl)); return o

    def relinearize(self, a, nl):
        o = np.zeros(2 * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_relinearize(self.h, _p(a), _p(o), nl)); return o

    def apply_galois(self, a, elt, nl):
        o = np.zeros(2 * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_apply_galois(self.h, _p(a), elt, _p(o), nl)); return o

    def rotate(self, a, steps, nl):
        o = np.zeros(2 * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_rotate(self.h, _p(a), steps, _p(o), nl)); return o

    def mod_switch_scale_to_next(self, a, pc, nl):
        o = np.zeros(pc * (nl - 1) * self.n, dtype=np.uint64); self._chk(self.R.rs_mod_switch_scale_to_next(self.h, _p(a), pc, _p(o), nl)); return o

    def drop_to_next(self, a, pc, nl):
        o = np.zeros(pc * (nl - 1) * self.n, dtype=np.uint64); self._chk(self.R.rs_drop_to_next(self.h, _p(a), pc, _p(o), nl)); return o

    def transform_to_ntt(self, a, pc, nl):
        o = np.zeros(pc * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_transform_to_ntt(self.h, _p(a), pc, _p(o), nl)); return o

    def transform_from_ntt(self, a, pc, nl):
        o = np.zeros(pc * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_transform_from_ntt(self.h, _p(a), pc, _p(o), nl)); return o

    def prng(self, kind, st, nl=0, nbytes=0):
        """kind 0: fill nbytes bytes; 1 ternary, 2 centered binomial, 3 uniform over nl limbs; st (np.uint64[4]) is advanced"""
        words = (nbytes + 7) // 8 if kind == 0 else nl * self.n
        o = np.zeros(words, dtype=np.uint64)
        self._chk(self.R.rs_prng(self.h, kind, _p(st), nl, nbytes, _p(o)))
        return o.view(np.uint8)[:nbytes] if kind == 0 else o

    def public_key(self):
        o = np.zeros(2 * self.nm * self.n, dtype=np.uint64); self._chk(self.R.rs_public_key(self.h, _p(o))); return o

    def encrypt_seeded(self, plain, asymmetric, nl, seed_lo, seed_hi=0, save_seed=False):
        o = np.zeros(2 * nl * self.n, dtype=np.uint64); sd = C.c_uint64(0)
        if plain is None:
            self._chk(self.R.rs_encrypt_seeded(self.h, None, 0, int(asymmetric), int(save_seed), nl, seed_lo, seed_hi, _p(o), C.byref(sd)))
        else:
            plain = np.ascontiguousarray(plain, dtype=np.uint64)
            self._chk(self.R.rs_encrypt_seeded(self.h, _p(plain), plain.size, int(asymmetric), int(save_seed), nl, seed_lo, seed_hi, _p(o), C.byref(sd)))
        return o, int(sd.value)

    def pack_rlwes(self, cts, nl, input_ntt, shift, input_interval, output_interval, apply_field_trace=True):
        flat = np.ascontiguousarray(np.concatenate(cts), dtype=np.uint64)
        o = np.zeros(2 * nl * self.n, dtype=np.uint64)
        self._chk(self.R.rs_pack_rlwes(self.h, _p(flat), len(cts), nl, int(input_ntt), shift, input_interval, output_interval, int(apply_field_trace), _p(o))); return o

    def extract_lwe(self, ct, term, nl, ntt_form):
        c0 = np.zeros(nl, dtype=np.uint64); c1 = np.zeros(nl * self.n, dtype=np.uint64)
        self._chk(self.R.rs_extract_lwe(self.h, _p(np.ascontiguousarray(ct, dtype=np.uint64)), term, nl, int(ntt_form), _p(c0), _p(c1))); return c0, c1

    def pack_lwes(self, lwes, nl, apply_field_trace=True):
        c0 = np.ascontiguousarray(np.concatenate([l[0] for l in lwes]), dtype=np.uint64); c1 = np.ascontiguousarray(np.concatenate([l[1] for l in lwes]), dtype=np.uint64)
        o = np.zeros(2 * nl * self.n, dtype=np.uint64)
        self._chk(self.R.rs_pack_lwes(self.h, _p(c0), _p(c1), len(lwes), nl, int(apply_field_trace), _p(o))); return o

    def plain_op(self, op, ct, nl, plain, correction_factor=1):
        """op 0 multiply_plain, 1 add_plain, 2 sub_plain with a plaintext modulo t; ct [2][nl][N] in the scheme's default form"""
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        o = np.zeros(2 * nl * self.n, dtype=np.uint64)
        self._chk(self.R.rs_plain_op(self.h, op, _p(np.ascontiguousarray(ct, dtype=np.uint64)), nl, _p(plain), plain.size, correction_factor, _p(o))); return o

    def batch_encode(self, values):
        values = np.ascontiguousarray(values, dtype=np.uint64); o = np.zeros(self.n, dtype=np.uint64)
        self._chk(self.R.rs_batch_encode(self.h, _p(values), values.size, _p(o))); return o

    def batch_decode(self, plain):
        plain = np.ascontiguousarray(plain, dtype=np.uint64); o = np.zeros(self.n, dtype=np.uint64)
        self._chk(self.R.rs_batch_decode(self.h, _p(plain), plain.size, _p(o))); return o

    def decrypt_raw(self, a, pc, nl, correction_factor=1):
        o = np.zeros((nl if self.scheme == 2 else 1) * self.n, dtype=np.uint64)
        self._chk(self.R.rs_decrypt_raw(self.h, _p(np.ascontiguousarray(a, dtype=np.uint64)), pc, nl, correction_factor, _p(o))); return o

    def decrypt_phase_raw(self, a, pc, nl):
        o = np.zeros(nl * self.n, dtype=np.uint64)
        self._chk(self.R.rs_decrypt_phase_raw(self.h, _p(np.ascontiguousarray(a, dtype=np.uint64)), pc, nl, _p(o))); return o

    def multiply_plain_raw(self, a, pc, plain, nl, ntt_form):
        o = np.zeros(pc * nl * self.n, dtype=np.uint64)
        self._chk(self.R.rs_multiply_plain_raw(self.h, _p(a), pc, _p(plain), _p(o), nl, int(ntt_form))); return o

    def galois_raw(self, a, pc, nl, elt, ntt_form):
        o = np.zeros(pc * nl * self.n, dtype=np.uint64); self._chk(self.R.rs_galois_raw(self.h, _p(a), pc, nl, elt, _p(o), int(ntt_form))); return o

    def time_op(self, op, nl, pc=1, warmup=1, iters=3):
        t = self.R.rs_time_op(self.h, op, nl, pc, warmup, iters)
        if t < 0:
            raise RuntimeError(self.R.rs_error(self.h).decode())
        return t
