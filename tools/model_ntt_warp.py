"""CPU model of the index logic of ntt_warp_kernel (csrc/b2_ntt_warp.cuh): shared-memory layouts, lane geometry of the
four round types, twiddle positions in the natural and in the row-major tables.  Small prime, plain integers; checks
the model against a layer-by-layer negacyclic NTT per SURVEY.md §9 C/D.  Developer tool (no GPU needed)."""
import numpy as np

LOGN, N = 16, 1 << 16
P = 786433            # 3 * 2^18 + 1


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1); x >>= 1
    return r


def find_psi():
    for g in range(2, 1000):
        psi = pow(g, (P - 1) // (2 * N), P)
        if pow(psi, N, P) == P - 1:
            return psi
    raise SystemExit("no root")


psi = find_psi(); ipsi = pow(psi, P - 2, P)
T = np.zeros(N, dtype=np.int64); IT = np.zeros(N, dtype=np.int64)
T[0] = 1; IT[N - 1] = 1
pw, ipw = psi, ipsi
for k in range(1, N):
    T[bitrev(k, LOGN)] = pw
    IT[bitrev(k - 1, LOGN)] = ipw          # device table = inv_root_powers shifted down by one
    pw = pw * psi % P; ipw = ipw * ipsi % P
INV_N = pow(N, P - 2, P)
INV_N_W = INV_N * int(IT[N - 2]) % P

# row-major copies (b2_ctx.cu)
TR = np.zeros(N, dtype=np.int64); ITR = np.zeros(N, dtype=np.int64)
A = LOGN - 8
for r in range(N >> 8):
    for s in range(8):
        for a in range(1 << s):
            TR[r * 256 + (1 << s) + a] = T[(1 << (A + s)) + r * (1 << s) + a]
            l = 7 - s
            ITR[r * 256 + (1 << s) + a] = IT[N - (N >> l) + r * (1 << s) + a]


def ref_fwd(x):
    x = x.copy()
    for l in range(LOGN):
        gap = 1 << (LOGN - l - 1); m = 1 << l
        i = np.arange(N // 2)
        r = T[m + i // gap]
        xi = 2 * gap * (i // gap) + (i % gap); yi = xi + gap
        X, Y = x[xi], x[yi] * r % P
        x[xi], x[yi] = (X + Y) % P, (X - Y) % P
    return x


def ref_inv(x):
    x = x.copy()
    for l in range(LOGN):
        gap = 1 << l; m = 1 << (LOGN - l - 1)
        i = np.arange(N // 2)
        r = IT[N - 2 * m + (i // gap)]      # shifted table: inv_root_powers[N - 2m + 1 + i/gap]
        xi = 2 * gap * (i // gap) + (i % gap); yi = xi + gap
        U, V = x[xi], x[yi]
        x[xi], x[yi] = (U + V) % P, (U - V) * r % P
    return x * INV_N % P


wl_col = lambda i: i ^ (((i >> 5) & 15) << 1)
wl_rs = lambda i: i ^ (((i >> 4) & 7) << 1)
wl_rx = lambda i: i ^ ((i >> 4) & 15)


def tw_global(fwd, XLO, l, j, xt, cnt):
    if fwd:
        b = (1 << l) + (xt >> (XLO + j + 1))
        return [int(T[b + a]) for a in range(cnt)]
    b = N - (N >> l) + (xt >> (l + 1))
    return [int(IT[b + a]) for a in range(cnt)]


def tw_staged(fwd, l, twr, c0, cnt):
    pos = (1 << (l - 8)) + (c0 >> (16 - l)) if fwd else (1 << (7 - l)) + (c0 >> (l + 1))
    return [int(twr[pos + a]) for a in range(cnt)]


def w_round(fwd, XLO, last, v, getter):
    for s in range(4):
        j = 3 - s if fwd else s
        l = 15 - (XLO + j) if fwd else XLO + j
        if not fwd and last and s == 3:
            for k in range(16):
                if k & (1 << j): continue
                U, V = v[k], v[k | (1 << j)]
                v[k], v[k | (1 << j)] = (U + V) * INV_N % P, (U - V) * INV_N_W % P
            continue
        w = getter(l, j, 16 >> (j + 1))
        for k in range(16):
            if k & (1 << j): continue
            a, b = k, k | (1 << j)
            if fwd:
                t = v[b] * w[k >> (j + 1)] % P
                v[a], v[b] = (v[a] + t) % P, (v[a] - t) % P
            else:
                U, V = v[a], v[b]
                v[a], v[b] = (U + V) % P, (U - V) * w[k >> (j + 1)] % P


def stage_col(limb, s):
    buf = [0] * 512
    for j in range(256):
        d = wl_col(2 * j)
        buf[d], buf[d + 1] = int(limb[(j << 8) + 2 * s]), int(limb[(j << 8) + 2 * s + 1])
    return buf


def stage_row(limb, s):
    buf = [0] * 512
    for j in range(256):
        d = wl_rs(2 * j)
        buf[d], buf[d + 1] = int(limb[s * 512 + 2 * j]), int(limb[s * 512 + 2 * j + 1])
    return buf


def slab(fwd, p2, src, dst, s):
    col_pass = (not p2) if fwd else p2
    if col_pass:
        buf = stage_col(src, s)
        regs = {}
        order = ("A", "B") if fwd else ("B", "A")
        for rnd_i, rnd in enumerate(order):
            newbuf = list(buf)
            for t in range(32):
                col, hi = t & 1, t >> 1
                if rnd == "A":
                    base, x0, xlo, sh = wl_col(t), ((t >> 1) << 8) + 2 * s + col, 12, 5
                else:
                    base, x0, xlo, sh = wl_col(col + (hi << 5)), ((hi << 4) << 8) + 2 * s + col, 8, 1
                idx = [base ^ wl_col(k << sh) for k in range(16)]
                v = [buf[i] for i in idx]
                w_round(fwd, xlo, (not fwd) and rnd == "A", v, lambda l, j, cnt: tw_global(fwd, xlo, l, j, x0, cnt))
                if rnd_i == 0:
                    for k in range(16): newbuf[idx[k]] = v[k]
                else:
                    for k in range(16): dst[x0 + (k << xlo)] = v[k]
            buf = newbuf
    else:
        buf = stage_row(src, s)
        twb = TR[s * 512:(s + 1) * 512] if fwd else ITR[s * 512:(s + 1) * 512]
        order = ("C", "D") if fwd else ("D", "C")
        for rnd_i, rnd in enumerate(order):
            newbuf = [None] * 512
            for t in range(32):
                lo, rl = t & 15, t >> 4
                twr = twb[rl * 256:(rl + 1) * 256]
                iC, iD = lo + (rl << 8), (lo << 4) + (rl << 8)
                if rnd == "C":
                    i0, xlo, sh, c0 = iC, 4, 4, lo
                else:
                    i0, xlo, sh, c0 = iD, 0, 0, lo << 4
                x0 = s * 512 + i0
                if rnd_i == 0:      # stage layout
                    if rnd == "C":
                        idx = [wl_rs(iC) ^ wl_rs(k << 4) for k in range(16)]
                    else:
                        sw = (lo & 7) << 1
                        idx = [iD + (k ^ sw) for k in range(16)]
                else:               # exchange layout
                    idx = [wl_rx(i0) ^ wl_rx(k << sh) for k in range(16)]
                v = [buf[i] for i in idx]
                w_round(fwd, xlo, False, v, lambda l, j, cnt: tw_staged(fwd, l, twr, c0, cnt))
                if rnd_i == 0:
                    for k in range(16): newbuf[wl_rx(i0) ^ wl_rx(k << sh)] = v[k]
                else:
                    for k in range(16): dst[x0 + (k << xlo)] = v[k]
            buf = newbuf


def model(fwd, x):
    mid = np.zeros(N, dtype=np.int64); out = np.zeros(N, dtype=np.int64)
    for s in range(128): slab(fwd, False, x, mid, s)
    for s in range(128): slab(fwd, True, mid, out, s)
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.integers(0, P, N, dtype=np.int64)
    f = ref_fwd(x)
    assert np.array_equal(ref_inv(f), x), "reference round trip"
    assert np.array_equal(model(True, x), f), "forward model mismatch"
    print("forward ok")
    assert np.array_equal(model(False, f), x), "inverse model mismatch"
    print("inverse ok")
