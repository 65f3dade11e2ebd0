"""CPU model of slzw_decode_fast_kernel's step algorithm (lzw_b200/csrc/decode_kernels.cu).

It mirrors the kernel's per-step logic (32 codes per step, classification, in-step length
resolution, word-start bit mask, source-pointer following, deferral rules) in plain Python, so
that the algorithm -- not the CUDA -- can be checked against the oracle without a GPU
(tests/test_fast_decode_model.py).  Test infrastructure, not part of the product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases as T    # noqa: E402

WIN = 1024
MAXSTACK = 4091
LIT = 1 << 31


def fast_decode(p, data: bytes, cap: int):
    """Returns (deferred, out bytes, status, detail)."""
    fixed = p.flavour == 1
    big = p.big_endian != 0
    inc = 1 if (not fixed and p.tiff_early_change) else 0
    cs = 8 if fixed else p.code_size
    if cap > (1 << 20):
        return True, b"", 0, 0
    roots = 256 if fixed else 1 << cs
    clear, eoi = roots, roots + 1
    first_index = 256 if fixed else roots + 2
    w = 12 if fixed else cs + 1
    mask = (1 << w) - inc
    nidx = first_index
    hp = False
    prev_off = prev_len = 0
    bitpos = 0
    total_bits = len(data) * 8
    out = bytearray()
    table = [0] * 4096
    padded = data + b"\0" * 8
    status = detail = 0

    def code_at(bit, w):
        ba = bit >> 3
        v = int.from_bytes(padded[ba:ba + 4], "little")
        sh = bit & 7
        if not big:
            return (v >> sh) & ((1 << w) - 1)
        vb = int.from_bytes(padded[ba:ba + 4], "big")
        return (vb >> (32 - w - sh)) & ((1 << w) - 1)

    while True:
        avail = (total_bits - bitpos) // w
        if avail == 0:
            if not fixed:
                status = 4
            break
        bmax = min(avail, 32)
        adj = 0 if hp else 1
        if not (fixed and nidx >= 4096):
            room = (mask if w < 12 else 4096) - nidx + adj
            bmax = min(bmax, room)
        bn = bmax if bmax else 1
        c = [code_at(bitpos + q * w, w) for q in range(bn)]
        b = bn
        ctrl = None
        if not fixed:
            for q in range(bn):
                if c[q] in (clear, eoi):
                    b = q
                    ctrl = c[q]
                    break
        if bmax == 0 and ctrl is None:
            return True, b"", 0, 0
        for q in range(adj, b):
            if c[q] >= roots and c[q] > nidx + q - adj:
                status, detail, b, ctrl = 2, c[q], q, None
                break
        if b > 0:
            produced = len(out)
            ln = [0] * b
            srci = [0] * b
            qd = [-2] * b
            for q in range(b):
                ni = nidx + q - adj
                if c[q] < roots:
                    ln[q] = 1
                    srci[q] = LIT | c[q]
                elif (not hp) and q == 0:
                    return True, b"", 0, 0
                elif c[q] < nidx:
                    e = table[c[q]]
                    ln[q] = e & 0xFFF
                    srci[q] = e >> 12
                else:
                    assert c[q] <= ni
                    qd[q] = c[q] - nidx + adj - 1
            known = [qd[q] < 0 for q in range(b)]
            for q in range(b):
                if qd[q] == -1:
                    ln[q] = prev_len + 1
            while not all(known):
                nl, nk = list(ln), list(known)
                for q in range(b):
                    frm = max(qd[q], 0)
                    if not known[q] and known[frm]:
                        nl[q] = ln[frm] + 1
                        nk[q] = True
                ln, known = nl, nk
            if any(l > MAXSTACK for l in ln):
                return True, b"", 0, 0
            incl = list(np.cumsum(ln))
            total = incl[b - 1]
            if total > WIN and b > 1:
                nb = sum(1 for q in range(b) if incl[q] <= WIN)
                b = max(nb, 1)
                ctrl = None
                status = detail = 0
                total = incl[b - 1]
            pos = [incl[q] - ln[q] for q in range(len(ln))]
            total_full = total
            if produced + total > cap:
                status, detail, total = 5, 0, cap - produced
            for q in range(b):
                if qd[q] >= 0:
                    srci[q] = produced + pos[qd[q]]
                elif qd[q] == -1:
                    srci[q] = prev_off
            for q in range(adj, b):
                idx = nidx + q - adj
                poff, plen = (prev_off, prev_len) if q == 0 else (produced + pos[q - 1], ln[q - 1])
                if idx < 4096:
                    table[idx] = (poff << 12) | ((plen + 1) & 0xFFF)
            # copy
            if total_full > WIN:
                so = srci[0]
                if so & LIT:
                    out.append(so & 0xFF)
                else:
                    dist = produced - so
                    snapshot = bytes(out)
                    for i in range(total):
                        out.append(snapshot[so + i % dist])
            else:
                bits = [0] * 32
                for q in range(b):
                    bits[pos[q] >> 5] |= 1 << (pos[q] & 31)
                nr = (total + 31) >> 5
                cnt = [0] * 32
                acc = 0
                for r in range(32):
                    cnt[r] = acc
                    acc += bin(bits[r]).count("1") if r < nr else 0
                snapshot = bytes(out)  # loads only touch bytes below `produced`
                step = bytearray(total)
                for ob in range(total):
                    r, l = ob >> 5, ob & 31
                    own = cnt[r] + bin(bits[r] & (0xFFFFFFFF >> (31 - l))).count("1") - 1
                    sp, i = srci[own], ob - pos[own]
                    while not (sp & LIT) and sp + i >= produced:
                        rel = sp + i - produced
                        assert rel < ob, "source pointers must strictly decrease"
                        own2 = cnt[rel >> 5] + bin(bits[rel >> 5] & (0xFFFFFFFF >> (31 - (rel & 31)))).count("1") - 1
                        sp, i = srci[own2], rel - pos[own2]
                    step[ob] = (sp & 0xFF) if (sp & LIT) else snapshot[sp + i]
                out += step
            prev_off = produced + pos[b - 1]
            prev_len = ln[b - 1]
            bitpos += b * w
            if not (fixed and nidx >= 4096):
                nidx += b - adj
                if not fixed and nidx == mask and w < 12:
                    w += 1
                    mask = (1 << w) - inc
            hp = True
        if status:
            break
        if ctrl is not None:
            bitpos += w
            if ctrl == eoi:
                break
            w = cs + 1
            mask = (1 << w) - inc
            nidx = first_index
            hp = False
    return False, bytes(out), status, detail


def check(p, raw: bytes, label: str, mutate=None, cap=None):
    st, _, packed = O.encode(p, raw)
    assert st == 0
    if mutate:
        packed = mutate(packed)
    cap = max(len(raw) * 2, 64) if cap is None else cap
    ost, odet, oout = O.decode(p, packed, cap=cap)
    deferred, out, st2, det2 = fast_decode(p, packed, cap=cap)
    if deferred:
        return "deferred(oracle status %d)" % ost
    assert (st2, det2) == (ost, odet), f"{label}: status {st2}/{det2} vs oracle {ost}/{odet}"
    assert out == oout, f"{label}: bytes differ"
    return "ok(status %d)" % ost


def main():
    rng = np.random.default_rng(7)
    res = {}
    for p in T.all_params():
        hi = T.max_symbol(p)
        for kind in T.KINDS:
            for n in (0, 1, 2, 5, 300, 5000, 40000):
                raw = T.make_stream(rng, kind, n, hi).tobytes()
                r = check(p, raw, f"{T.pname(p)} {kind} {n}")
                res[r] = res.get(r, 0) + 1
                for cap in (n, n // 2, max(n - 1, 0)):
                    r = check(p, raw, f"{T.pname(p)} {kind} {n} cap {cap}", cap=cap)
                    res[r] = res.get(r, 0) + 1
    # corrupt streams: the model must either defer or agree with the oracle
    for p in (O.gif(8), O.tiff(), O.gif(3), O.fixed(False), O.fixed(True)):
        hi = T.max_symbol(p)
        for t in range(200):
            raw = T.make_stream(rng, T.KINDS[t % len(T.KINDS)], int(rng.integers(10, 3000)), hi).tobytes()
            def mut(pk):
                a = bytearray(pk)
                for _ in range(int(rng.integers(1, 4))):
                    if a:
                        a[int(rng.integers(0, len(a)))] ^= 1 << int(rng.integers(0, 8))
                if rng.random() < 0.3 and len(a) > 2:
                    del a[int(rng.integers(1, len(a))):]
                return bytes(a)
            r = check(p, raw, f"{T.pname(p)} corrupt {t}", mutate=mut, cap=len(raw) if t % 2 else None)
            res[r] = res.get(r, 0) + 1
    print(res)


if __name__ == "__main__":
    main()


def long_words():
    """Words longer than the step window (the periodic-copy path)."""
    for p in (O.tiff(), O.gif(2), O.fixed(False)):
        raw = bytes(500_000) if p.flavour == 0 else bytes(200_000)
        print(T.pname(p), "zeros", check(p, raw, "zeros"))
        raw = (b"ab" * 200_000)
        if p.code_size == 2 and p.flavour == 0:
            raw = bytes([1, 2]) * 200_000
        print(T.pname(p), "period2", check(p, raw, "period2"))
