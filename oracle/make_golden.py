"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, host branch).

Run in the build container only (needs /root/reference to build oracle/_ref):
    python oracle/make_golden.py
Each file pins, for one scheme at N=256 with CoeffModulus::create(256, {50,40,40,50}) and context seed 0x123:
moduli, plain modulus, psi roots, relin + Galois keys, a fresh encryption of zero and the reference's outputs of
multiply, square, relinearize, rotate(1), mod_switch/rescale_to_next (two levels), NTT / INTT in all three
table-indexer modes, the Galois maps and multiply_plain — the raw uint64 words the reference produced.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pyoracle as po

N, LOGN, BITS = 256, 8, [50, 40, 40, 50]
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        r = po.Reference(scheme, N, BITS, t_bits=20, seed=0x123)
        K = r.nm; L = K - 1
        rng = np.random.default_rng(1000 + scheme)
        g = {"n": N, "scheme": scheme, "moduli": r.moduli, "t": np.uint64(r.t), "bits": np.array(BITS)}
        roots = []
        for i in range(K):
            root, rp, irp = r.ntt_table(i, N)
            roots.append(root); g["root_powers_%d" % i] = rp; g["inv_root_powers_%d" % i] = irp
        g["roots"] = np.array(roots, dtype=np.uint64)
        r.gen_relin(); gal = r.galois_element_from_step(1); r.gen_galois([gal, 2 * N - 1])
        g["galois_elt"] = np.uint64(gal)
        g["relin_key"] = r.relin_key(); g["galois_key"] = r.galois_key(gal); g["secret_key"] = r.secret_key()
        ct = r.encrypt_zero(); g["ct"] = ct
        ntt = scheme != po.BFV

        def rnd(pc, nl):
            out = np.empty((pc, nl, N), dtype=np.uint64)
            for j in range(nl):
                out[:, j, :] = rng.integers(0, int(r.moduli[j]), size=(pc, N), dtype=np.uint64)
            return out.reshape(-1)
        for lvl in (L, L - 1):
            a, b, c3 = rnd(2, lvl), rnd(2, lvl), rnd(3, lvl)
            t = "_l%d" % lvl
            g["a" + t], g["b" + t], g["c3" + t] = a, b, c3
            g["mul" + t] = r.multiply(a, b, lvl); g["square" + t] = r.square(a, lvl)
            g["relin" + t] = r.relinearize(c3, lvl); g["rotate1" + t] = r.rotate(a, 1, lvl)
            g["modswitch" + t] = r.mod_switch_scale_to_next(a, 2, lvl)
            if scheme == po.CKKS:
                g["drop" + t] = r.drop_to_next(a, 2, lvl)
            g["galois_coeff" + t] = r.galois_raw(a, 2, lvl, gal, False); g["galois_ntt" + t] = r.galois_raw(a, 2, lvl, gal, True)
            pl = rnd(1, lvl); g["plain" + t] = pl
            g["mulplain_ntt" + t] = r.multiply_plain_raw(a, 2, pl, lvl, True) if ntt else r.multiply_plain_raw(r.transform_to_ntt(a, 2, lvl), 2, pl, lvl, True)
            g["to_ntt" + t] = r.transform_to_ntt(a, 2, lvl); g["from_ntt" + t] = r.transform_from_ntt(a, 2, lvl)
        x = rng.integers(0, 1 << 39, size=(L + 1) * L * N, dtype=np.uint64)
        g["ks_set_in"] = x; g["ks_set_out"] = r.ntt_fwd(x, L + 1, L, 1, 0, L)
        y = rng.integers(0, 1 << 39, size=2 * (L + 1) * N, dtype=np.uint64)
        g["ks_skip_in"] = y; g["ks_skip_out"] = r.ntt_inv(y, 2, L + 1, 2, 0, L)
        np.savez_compressed(os.path.join(OUT, "ref_%s_n256.npz" % name), **g)
        print("wrote", name, {k: getattr(v, "shape", None) for k, v in list(g.items())[:4]})


def main_decrypt():
    """tests/golden/ref_decrypt_n256.npz: the reference's Decryptor (src/decryptor.cu) on random ciphertexts of two and three
    polynomials at the first and the last level, under the reference's own secret key; BGV also with a correction factor."""
    g = {"n": N, "bits": np.array(BITS)}
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        r = po.Reference(scheme, N, BITS, t_bits=20, seed=0x123)
        L = r.nm - 1
        rng = np.random.default_rng(2000 + scheme)
        g[name + "_moduli"] = r.moduli; g[name + "_t"] = np.uint64(r.t); g[name + "_sk"] = r.secret_key()
        zero = r.encrypt_zero()
        g[name + "_zero"] = zero; g[name + "_zero_dec"] = r.decrypt_raw(zero, 2, L)
        for nl in (L, 1):
            for pc in (2, 3):
                ct = np.empty((pc, nl, N), dtype=np.uint64)
                for j in range(nl):
                    ct[:, j, :] = rng.integers(0, int(r.moduli[j]), size=(pc, N), dtype=np.uint64)
                ct = ct.reshape(-1)
                key = "%s_l%d_p%d" % (name, nl, pc)
                g[key + "_ct"] = ct; g[key + "_dec"] = r.decrypt_raw(ct, pc, nl)
                if scheme == po.BGV:
                    g[key + "_dec_cf7"] = r.decrypt_raw(ct, pc, nl, 7)
                if scheme == po.BFV:
                    g[key + "_phase"] = r.decrypt_phase_raw(ct, pc, nl)
    np.savez_compressed(os.path.join(OUT, "ref_decrypt_n256.npz"), **g)
    print("wrote decrypt", len(g), "arrays")


def main_encrypt():
    """tests/golden/ref_encrypt_n256.npz: the reference's generator, samplers and Encryptor (src/utils/random_generator.cu, src/utils/rlwe.cu,
    src/encryptor.cu) with explicit seeds: AES-CTR bytes, the three samplers, zero encryptions (asymmetric / symmetric, with and without a
    saved seed) at the key level, the first level and the last level, and encryptions of a message, under the reference's own keys."""
    g = {"n": N, "bits": np.array(BITS)}
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        r = po.Reference(scheme, N, BITS, t_bits=20, seed=0x123)
        K = r.nm; L = K - 1
        rng = np.random.default_rng(3000 + scheme)
        g[name + "_moduli"] = r.moduli; g[name + "_t"] = np.uint64(r.t); g[name + "_sk"] = r.secret_key(); g[name + "_pk"] = r.public_key()
        st = np.array([77, 3, 9, 0], dtype=np.uint64)
        g[name + "_bytes"] = r.prng(0, st, nbytes=16 * 7 + 3)
        for kind, kn in ((1, "ternary"), (2, "cbd"), (3, "uniform")):
            g["%s_%s" % (name, kn)] = r.prng(kind, st, nl=K)
        g[name + "_state_after"] = st.copy()
        for nl in (K, L, 1):
            g["%s_l%d_asym" % (name, nl)] = r.encrypt_seeded(None, True, nl, 1234, 5)[0]
            for save in (0, 1):
                ct, seed = r.encrypt_seeded(None, False, nl, 99, 0, save_seed=bool(save))
                g["%s_l%d_sym%d" % (name, nl, save)] = ct; g["%s_l%d_sym%d_seed" % (name, nl, save)] = np.uint64(seed)
        if scheme == po.CKKS:
            plain = np.concatenate([rng.integers(0, int(r.moduli[j]), size=N, dtype=np.uint64) for j in range(L)])
        else:
            plain = rng.integers(0, r.t, size=N // 2 + 3, dtype=np.uint64)
        g[name + "_plain"] = plain
        for asym in (0, 1):
            g["%s_enc%d" % (name, asym)] = r.encrypt_seeded(plain, bool(asym), L, 4242)[0]
    np.savez_compressed(os.path.join(OUT, "ref_encrypt_n256.npz"), **g)
    print("wrote encrypt", len(g), "arrays")


def main_keygen():
    """tests/golden/ref_keygen_n256.npz: the reference's KeyGenerator from a context seed (src/key_generator.cu): secret key, public key, relinearisation key and
    two Galois keys, in the order they were drawn from the context's generator."""
    g = {"n": N, "bits": np.array(BITS), "seed": np.uint64(0xabcdef12345)}
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        r = po.Reference(scheme, N, BITS, t_bits=20, seed=0xabcdef12345)
        r.gen_relin(); r.gen_galois([3, 2 * N - 1])
        g[name + "_moduli"] = r.moduli; g[name + "_t"] = np.uint64(r.t)
        g[name + "_sk"] = r.secret_key(); g[name + "_pk"] = r.public_key(); g[name + "_relin"] = r.relin_key()
        g[name + "_galois_3"] = r.galois_key(3); g[name + "_galois_conj"] = r.galois_key(2 * N - 1)
    np.savez_compressed(os.path.join(OUT, "ref_keygen_n256.npz"), **g)
    print("wrote keygen", len(g), "arrays")


PACK_CASES = ((4, 0, 64, 16, 1), (3, 0, 64, 16, 1), (8, 5, 64, 8, 0), (2, 0, 32, 16, 1), (1, 0, 64, 64, 1), (5, 3, 64, 8, 1))


def main_pack():
    """tests/golden/ref_pack_n64.npz: Evaluator::pack_rlwe_ciphertexts_new (src/evaluator_lwes.cu:305-491) of the unmodified reference at N = 64,
    moduli {40, 40, 50}: random input ciphertexts, the Galois keys it used, its outputs (count, shift, input_interval, output_interval, trace)."""
    n = 64
    g = {"n": n, "bits": np.array([40, 40, 50]), "cases": np.array(PACK_CASES)}
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        r = po.Reference(scheme, n, [40, 40, 50], t_bits=20, seed=0x123)
        L = r.nm - 1
        elts = sorted({(n >> k) + 1 for k in range(0, 6)} | {2 ** k + 1 for k in range(1, 7)})
        elts = [e for e in elts if e < 2 * n]
        r.gen_galois(elts)
        g[name + "_moduli"] = r.moduli; g[name + "_t"] = np.uint64(r.t); g[name + "_elts"] = np.array(elts, dtype=np.uint64)
        for e in elts:
            g["%s_key_%d" % (name, e)] = r.galois_key(e)
        rng = np.random.default_rng(4000 + scheme)
        for ci, (count, shift, ii, oi, ft) in enumerate(PACK_CASES):
            cts = []
            for _ in range(count):
                c = np.empty((2, L, n), dtype=np.uint64)
                for j in range(L):
                    c[:, j, :] = rng.integers(0, int(r.moduli[j]), size=(2, n), dtype=np.uint64)
                cts.append(c.reshape(-1))
            g["%s_c%d_in" % (name, ci)] = np.concatenate(cts)
            for inp_ntt in ((0, 1) if scheme != po.BFV else (0,)):
                g["%s_c%d_out%d" % (name, ci, inp_ntt)] = r.pack_rlwes(cts, L, bool(inp_ntt), shift, ii, oi, bool(ft))
    np.savez_compressed(os.path.join(OUT, "ref_pack_n64.npz"), **g)
    print("wrote pack", len(g), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pack":
        main_pack()
    elif len(sys.argv) > 1 and sys.argv[1] == "keygen":
        main_keygen()
    elif len(sys.argv) > 1 and sys.argv[1] == "decrypt":
        main_decrypt()
    elif len(sys.argv) > 1 and sys.argv[1] == "encrypt":
        main_encrypt()
    else:
        main(); main_decrypt(); main_encrypt(); main_pack(); main_keygen()
