#!/usr/bin/env python
"""numpy model of the index conventions of csrc/sparc_pair.cuh (development aid, CPU only).

Emulates, thread by thread, (i) the pass structure the pair kernels had until round 2 (separate real-FFT
pre / post-processing passes) and (ii) the fused forms: the forward transform's last radix-4 pass that also
does the post-processing of the mirror pairs (k, H-k), and the transposed transform's first radix-4 pass that
builds its inputs from the row block (plan 4,8,8,8).  Checks both against the dense DCT matrix and counts
the shared-memory bank conflicts of the fused accesses under the slot swizzles.

usage: tools/model_pair_fused.py
"""
import numpy as np

LOGH = 11
H = 1 << LOGH
N = 2 * H
Hh = H // 2
Q = H // 4
RS = 0.70710678118654752440


def brev(k, bits):
    r = 0
    for i in range(bits):
        r |= ((k >> i) & 1) << (bits - 1 - i)
    return r


def pswz_fwd(i):      # pswz<11>
    return i ^ ((i >> 3) & 3) ^ (((i >> 5) & 1) << 2)


def pswz_inv(i):      # swizzle of the transposed plan (4,8,8,8): bits 3,4,5 -> 0,1,2
    return i ^ ((i >> 3) & 7)


def radix_dft(v, logr, inv):
    """in place on a python list of complex; output m holds frequency brev(m) (as radix_dft<> in the .cuh)"""
    mi = (1j if inv else -1j)
    w8 = np.exp((1j if inv else -1j) * np.pi / 4)

    def bf(a, b):
        v[a], v[b] = v[a] + v[b], v[a] - v[b]
    if logr == 2:
        bf(0, 2); bf(1, 3)
        v[3] *= mi
        bf(0, 1); bf(2, 3)
    else:
        bf(0, 4); bf(1, 5); bf(2, 6); bf(3, 7)
        v[5] *= w8; v[6] *= mi; v[7] *= mi * w8
        bf(0, 2); bf(1, 3); bf(4, 6); bf(5, 7)
        v[3] *= mi; v[7] *= mi
        bf(0, 1); bf(2, 3); bf(4, 5); bf(6, 7)


def twiddle(v, logr, logs, j, inv):
    S = 1 << logs
    sgn = 1j if inv else -1j
    for m in range(1 << logr):
        v[m] *= np.exp(sgn * 2 * np.pi * j * brev(m, logr) / S)


def dif_pass(buf, logs, logr, inv):
    logq = logs - logr
    for q in range(H >> logr):
        j = q & ((1 << logq) - 1)
        base = ((q >> logq) << logs) | j
        v = [buf[base + (m << logq)] for m in range(1 << logr)]
        radix_dft(v, logr, inv)
        if logq > 0:
            twiddle(v, logr, logs, j, inv)
        for m in range(1 << logr):
            buf[base + (m << logq)] = v[m]


def hp_of(k):
    return 0.5 * np.exp(-1j * np.pi * k / (2 * N))


def mirror(hp):
    return np.exp(-1j * np.pi / 4) * np.conj(hp)


def post_pair(za, zb, hp):
    """returns the new contents of the slots of Z[k] and Z[H-k]: (Y[k], Y[N-k]) and (Y[H-k], Y[H+k]) as re + i im"""
    q = (2 * hp) ** 4
    E2 = za + np.conj(zb)
    O2 = -1j * (za - np.conj(zb))
    qO = O2 * q
    ta = (E2 + qO) * hp
    t2 = np.conj(E2 - qO) * mirror(hp)
    return np.conj(ta), np.conj(t2)


def pre_pair(ca, cb, hp):
    """ca = c[k] + i c[N-k], cb = c[H-k] + i c[H+k]  ->  Z[k], Z[H-k]"""
    q = (2 * hp) ** 4
    U = ca * hp
    V = cb * mirror(hp)
    A = complex(U.real + V.real, V.imag - U.imag)
    D = complex(U.real - V.real, -(U.imag + V.imag))
    B = D * np.conj(q)
    return complex(A.real - B.imag, A.imag + B.real), complex(A.real + B.imag, B.real - A.imag)


def makhoul(p):
    return N - 1 - (p - 1) // 2 if p & 1 else p // 2


def conflicts(slot_lists):
    """slot_lists[lane] = slot accessed by the lane in one 128-bit warp access; served 8 lanes at a time"""
    extra = 0
    for g in range(0, len(slot_lists), 8):
        grp = [s & 7 for s in slot_lists[g:g + 8]]
        extra += max(np.bincount(grp, minlength=8)) - 1
    return extra


def forward_unfused(x_full):
    v = np.zeros(N)
    for p in range(N):
        v[makhoul(p)] = x_full[p]
    buf = (v[0::2] + 1j * v[1::2]).astype(complex)
    for logs, logr in ((11, 3), (8, 3), (5, 3), (2, 2)):
        dif_pass(buf, logs, logr, False)
    for k in range(Hh + 1):
        pa, pb = brev(k, LOGH), brev((H - k) & (H - 1), LOGH)
        if k == 0:
            z = buf[0]
            buf[0] = complex(z.real + z.imag, (z.real - z.imag) * RS)
        else:
            a, b = post_pair(buf[pa], buf[pb], hp_of(k))
            buf[pa] = a
            if pb != pa:
                buf[pb] = b
    return buf


def forward_fused(x_full):
    v = np.zeros(N)
    for p in range(N):
        v[makhoul(p)] = x_full[p]
    buf = (v[0::2] + 1j * v[1::2]).astype(complex)
    for logs, logr in ((11, 3), (8, 3), (5, 3)):
        dif_pass(buf, logs, logr, False)
    lanes_v = [[] for _ in range(4)]
    lanes_u = [[] for _ in range(4)]
    for tk in range(Q // 2):
        if tk == 0:
            q1, q2 = 0, 1
        else:
            q1 = 2 * tk
            b = brev(q1, LOGH - 2)
            assert 1 <= b < Q // 2
            bp = Q - b
            q2 = brev(bp, LOGH - 2)
        vv = [buf[4 * q1 + m] for m in range(4)]
        uu = [buf[4 * q2 + m] for m in range(4)]
        for m in range(4):
            lanes_v[m].append(pswz_fwd(4 * q1 + m))
            lanes_u[m].append(pswz_fwd(4 * q2 + m))
        radix_dft(vv, 2, False)
        radix_dft(uu, 2, False)
        if tk == 0:
            z = vv[0]
            vv[0] = complex(z.real + z.imag, (z.real - z.imag) * RS)
            vv[1], _ = post_pair(vv[1], vv[1], hp_of(Hh))
            vv[2], vv[3] = post_pair(vv[2], vv[3], hp_of(Q))
            uu[0], uu[3] = post_pair(uu[0], uu[3], hp_of(Q // 2))
            uu[2], uu[1] = post_pair(uu[2], uu[1], hp_of(3 * Q // 2))
        else:
            vv[0], uu[3] = post_pair(vv[0], uu[3], hp_of(b))
            vv[2], uu[1] = post_pair(vv[2], uu[1], hp_of(b + Q))
            uu[0], vv[3] = post_pair(uu[0], vv[3], hp_of(bp))
            uu[2], vv[1] = post_pair(uu[2], vv[1], hp_of(bp + Q))
        for m in range(4):
            buf[4 * q1 + m] = vv[m]
            buf[4 * q2 + m] = uu[m]
    extra = sum(conflicts(l) for l in lanes_v) + sum(conflicts(l) for l in lanes_u)
    total = 8 * (Q // 2) // 8
    return buf, extra, total


def transposed_unfused(c_full):
    """c_full[K], K = 0..N-1 (zero where the block has no row); returns the buffer after the inverse FFT"""
    buf = np.zeros(H, complex)
    for k in range(Hh + 1):
        if k == 0:
            wH = c_full[H] * RS
            buf[0] = complex(c_full[0] + wH, c_full[0] - wH)
            continue
        ca = complex(c_full[k], c_full[N - k])
        cb = complex(c_full[H - k], c_full[H + k])
        a, b = pre_pair(ca, cb, hp_of(k))
        buf[k] = a
        if k != Hh:
            buf[H - k] = b
    for logs, logr in ((11, 3), (8, 3), (5, 3), (2, 2)):
        dif_pass(buf, logs, logr, True)
    return buf


def transposed_fused(c_full):
    buf = np.zeros(H, complex)

    def pp(k):
        return pre_pair(complex(c_full[k], c_full[N - k]), complex(c_full[H - k], c_full[H + k]), hp_of(k))
    lanes = [[] for _ in range(8)]
    for tk in range(Q // 2):
        vv, uu = [0j] * 4, [0j] * 4
        if tk == 0:
            j, jp = 0, Q // 2
            wH = c_full[H] * RS
            vv[0] = complex(c_full[0] + wH, c_full[0] - wH)
            vv[2], _ = pp(Hh)
            vv[1], vv[3] = pp(Q)
            uu[0], uu[3] = pp(Q // 2)
            uu[1], uu[2] = pp(3 * Q // 2)
        else:
            j, jp = tk, Q - tk
            vv[0], uu[3] = pp(j)
            vv[1], uu[2] = pp(j + Q)
            uu[0], vv[3] = pp(jp)
            uu[1], vv[2] = pp(jp + Q)
        radix_dft(vv, 2, True)
        radix_dft(uu, 2, True)
        twiddle(vv, 2, LOGH, j, True)
        twiddle(uu, 2, LOGH, jp, True)
        for m in range(4):
            buf[j + Q * m] = vv[m]
            buf[jp + Q * m] = uu[m]
            lanes[m].append(pswz_inv(j + Q * m))
            lanes[4 + m].append(pswz_inv(jp + Q * m))
    for logs, logr in ((9, 3), (6, 3), (3, 3)):
        dif_pass(buf, logs, logr, True)
    extra = sum(conflicts(l) for l in lanes)
    return buf, extra


def pass_conflicts(logs, logr, swz):
    logq = logs - logr
    extra = 0
    for m in range(1 << logr):
        lane_slots = []
        for q in range(H >> logr):
            j = q & ((1 << logq) - 1)
            base = ((q >> logq) << logs) | j
            lane_slots.append(swz(base + (m << logq)))
        extra += conflicts(lane_slots)
    return extra


def main():
    rng = np.random.RandomState(1)
    Mr, Mc = 3413, 1024
    o0 = rng.permutation(np.arange(1, N))[:Mr]
    o1 = rng.permutation(np.arange(1, N))[:Mc]
    x = rng.randn(Mc)
    z = rng.randn(Mr)
    kk, mm = np.meshgrid(o0, o1, indexing="ij")
    D = np.cos(np.pi * kk * (2 * mm + 1) / (2.0 * N))          # [Mr, Mc]
    # forward
    x_full = np.zeros(N)
    x_full[o1] = x
    bu = forward_unfused(x_full)
    bf, extra, total = forward_fused(x_full)

    def read_fwd(buf):
        out = np.zeros(Mr)
        for i, K in enumerate(o0):
            if K < H:
                out[i] = buf[brev(K, LOGH)].real
            elif K == H:
                out[i] = buf[0].imag
            else:
                out[i] = buf[brev(N - K, LOGH)].imag
        return out
    yu, yf, yd = read_fwd(bu), read_fwd(bf), D @ x
    res = {}
    su = (yu @ yd) / (yd @ yd)
    res.update(fwd_scale=su, fwd_err_dense=np.abs(yu / su - yd).max(), fwd_err_fused=np.abs(bf - bu).max(), fwd_conflicts=extra)
    print("forward: unfused vs dense (scale %.6f) max err %.3e; fused vs unfused max err %.3e; fused-pass extra conflict wavefronts %d"
          % (su, np.abs(yu / su - yd).max(), np.abs(bf - bu).max(), extra))
    # transposed
    c_full = np.zeros(N + 1)
    c_full[o0] = z
    tu = transposed_unfused(c_full)
    tf, extra_t = transposed_fused(c_full)

    def read_tr(buf):
        out = np.zeros(Mc)
        for j, K in enumerate(o1):
            m = makhoul(int(K))
            e = buf[brev(m >> 1, LOGH)]
            out[j] = e.imag if m & 1 else e.real
        return out
    xu, xf, xd = read_tr(tu), read_tr(tf), D.T @ z
    st = (xu @ xd) / (xd @ xd)
    res.update(tr_scale=st, tr_err_dense=np.abs(xu / st - xd).max(), tr_err_fused=np.abs(tf - tu).max(), tr_conflicts=extra_t)
    print("transposed: unfused vs dense (scale %.6f) max err %.3e; fused vs unfused max err %.3e; fused-pass extra conflict wavefronts %d"
          % (st, np.abs(xu / st - xd).max(), np.abs(tf - tu).max(), extra_t))
    for name, plan, swz in (("forward (8,8,8,4) pswz<11>", ((11, 3), (8, 3), (5, 3), (2, 2)), pswz_fwd),
                            ("transposed (4,8,8,8) i^((i>>3)&7)", ((11, 2), (9, 3), (6, 3), (3, 3)), pswz_inv)):
        pc = [int(pass_conflicts(s, r, swz)) for s, r in plan]
        res.setdefault("pass_conflicts", []).append(pc)
        print(name, "extra conflict wavefronts per pass:", pc)
    return res


if __name__ == "__main__":
    main()
