"""Lane-level Python model of lzw_b200/csrc/predictor_kernels.cu (test infrastructure).

The two TIFF predictor kernels are index arithmetic on aligned 16-byte words: ragged first / last
words, row starts inside a word, byte masks built from nibbles, a segmented warp scan by
subtraction in 16-bit fields, a byte rotation for 3 samples per pixel.  This model restates that
arithmetic step by step (same names as the kernel) so that it can be checked against the numpy
restatement of TIFF 6.0 section 14 without a GPU; tests/test_predictor_model.py runs it."""
import numpy as np

M32 = 0xFFFFFFFF
FIELDS = 0x00FF00FF


def vadd4(a, b):
    return sum((((a >> (8 * i)) + (b >> (8 * i))) & 0xFF) << (8 * i) for i in range(4))


def vsub4(a, b):
    return sum((((a >> (8 * i)) - (b >> (8 * i))) & 0xFF) << (8 * i) for i in range(4))


def byte_perm(x, y, s):
    src = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    return sum(src[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def funnelshift_l(lo, hi, sh):
    sh &= 31
    return ((hi << sh) | (lo >> (32 - sh))) & M32 if sh else hi


def nibble_to_bytes(nib):
    return ((((nib & 0xF) * 0x00204081) & 0x01010101) * 0xFF) & M32


def load_word(mem, p, jlo, jhi):
    w = [0] * 4
    for j in range(16):
        if jlo <= j < jhi:
            w[j >> 2] |= int(mem[p + j]) << (8 * (j & 3))
    return w


def store_word(mem, p, w, jlo, jhi):
    for j in range(16):
        if jlo <= j < jhi:
            mem[p + j] = (w[j >> 2] >> (8 * (j & 3))) & 0xFF


def hdiff_strip(mem, addr, length, row_bytes, spp, threads=8):
    """slzw_hdiff_kernel for one strip at byte address `addr` of `mem`, in place."""
    if length == 0:
        return
    piece = threads * 16
    step = piece % row_bytes
    skew = addr & 15
    base = addr - skew
    nwords = (skew + length + 15) >> 4
    pieces = (nwords + threads - 1) // threads
    wis = [(pieces - 1) * threads + t for t in range(threads)]
    cols = [((wi * 16 - skew) % row_bytes if wi * 16 >= skew else 0) for wi in wis]
    for _ in range(pieces):
        pending = []
        for t in range(threads):
            wi, col = wis[t], cols[t]
            if wi < nwords:
                w16 = wi * 16
                jlo = 0 if w16 >= skew else skew - w16
                jhi = 16 if w16 + 16 <= skew + length else skew + length - w16
                v = load_word(mem, base + w16, jlo, jhi)
                prev = 0
                for j in range(4):
                    if w16 + j >= skew + 4:
                        prev |= int(mem[base + w16 - 4 + j]) << (8 * j)
                keep = 0
                if jlo > 0:
                    j0 = jlo
                else:
                    j0 = 0 if col == 0 else row_bytes - col
                    if col < spp:
                        keep = (1 << (spp - col)) - 1
                while j0 < 16:
                    keep |= ((1 << spp) - 1) << j0
                    j0 += row_bytes
                x = [prev] + v
                d = [0] * 4
                for k in range(4):
                    left = x[k] if spp == 4 else funnelshift_l(x[k], x[k + 1], 8 * spp)
                    m = nibble_to_bytes(keep >> (4 * k))
                    d[k] = (vsub4(v[k], left) & ~m & M32) | (v[k] & m)
                pending.append((base + w16, d, jlo, jhi))
            cols[t] = col - step if col >= step else col + row_bytes - step
            wis[t] -= threads
        for a in pending:  # after the barrier
            store_word(mem, *a)


def accumulate_rows(mem, addr, length, row_bytes, spp, lane_words=2):
    """accumulate_rows<SPP, W> for one warp's group of rows, in place."""
    lane_bytes = 16 * lane_words
    pass_bytes = 32 * lane_bytes
    skew = addr & 15
    base = addr - skew
    nwords = (skew + length + 15) >> 4
    step = pass_bytes % row_bytes
    cols = [(lane * lane_bytes + row_bytes * 16 - skew) % row_bytes for lane in range(32)]
    phases = [(lane * lane_bytes + 48 - skew) % 3 for lane in range(32)]
    carry_lo = carry_hi = 0
    for w0 in range(0, nwords, 32 * lane_words):
        lanes = []
        for lane in range(32):
            wi = w0 + lane * lane_words
            vs, jl, jh = [], [], []
            for k in range(lane_words):
                w, w16 = wi + k, (wi + k) * 16
                jlo = jhi = 0
                v = [0] * 4
                if w < nwords:
                    jlo = 0 if w16 >= skew else skew - w16
                    jhi = 16 if w16 + 16 <= skew + length else skew + length - w16
                    v = load_word(mem, base + w16, jlo, jhi)
                vs.append(v), jl.append(jlo), jh.append(jhi)
            col = cols[lane]
            j0 = 0 if col == 0 else row_bytes - col
            starts = 0
            while j0 < lane_bytes:
                starts |= 1 << j0
                j0 += row_bytes
            r = [0] * spp
            o = [[0] * 4 for _ in range(lane_words)]
            for j in range(lane_bytes):
                if (starts >> j) & 1:
                    r = [0] * spp
                r[j % spp] += (vs[j >> 4][(j >> 2) & 3] >> (8 * (j & 3))) & 0xFF
                o[j >> 4][(j >> 2) & 3] |= (r[j % spp] & 0xFF) << (8 * (j & 3))
            tot = sum((r[c] & 0xFF) << (8 * c) for c in range(spp))
            to_rel = 0x3210
            if spp == 3:
                phase = phases[lane]
                tot = byte_perm(tot, 0, (0x3210, 0x3102, 0x3021)[phase])
                to_rel = (0x3210, 0x3021, 0x3102)[phase]
            lanes.append(dict(wi=wi, jl=jl, jh=jh, o=o, tot=tot, to_rel=to_rel, starts=starts))
            c = col + step
            cols[lane] = c - row_bytes if c >= row_bytes else c
            phases[lane] = (phases[lane] + pass_bytes % 3) % 3
        tot_lo = [ln["tot"] & FIELDS for ln in lanes]
        tot_hi = [(ln["tot"] >> 8) & FIELDS for ln in lanes]
        lo, hi = list(tot_lo), list(tot_hi)
        d = 1
        while d < 32:
            lo = [(lo[i] + lo[i - d]) & M32 if i >= d else lo[i] for i in range(32)]
            hi = [(hi[i] + hi[i - d]) & M32 if i >= d else hi[i] for i in range(32)]
            d <<= 1
        excl_lo = [(lo[i] - tot_lo[i]) & M32 for i in range(32)]
        excl_hi = [(hi[i] - tot_hi[i]) & M32 for i in range(32)]
        started = sum((1 << i) for i in range(32) if lanes[i]["starts"])
        for lane, ln in enumerate(lanes):
            below = started & ((1 << lane) - 1)
            if below:
                seg = below.bit_length() - 1
                b_lo, b_hi = (excl_lo[lane] - excl_lo[seg]) & M32, (excl_hi[lane] - excl_hi[seg]) & M32
            else:
                b_lo, b_hi = (excl_lo[lane] + carry_lo) & M32, (excl_hi[lane] + carry_hi) & M32
            before = (b_lo & FIELDS) | ((b_hi & FIELDS) << 8)
            if spp == 3:
                before = byte_perm(before, 0, ln["to_rel"])
            if ln["wi"] < nwords:
                st = ln["starts"]
                ahead = ((1 << ((st & -st).bit_length() - 1)) - 1) if st else M32
                for k in range(lane_words):
                    o = ln["o"][k]
                    for q in range(4):
                        kk = 4 * k + q
                        sel = sum(((4 * kk + i) % spp) << (4 * i) for i in range(4))
                        add = byte_perm(before, 0, sel)
                        if st:
                            add &= nibble_to_bytes(ahead >> (16 * k + 4 * q))
                        o[q] = vadd4(o[q], add)
                    if ln["wi"] + k < nwords:
                        store_word(mem, base + (ln["wi"] + k) * 16, o, ln["jl"][k], ln["jh"][k])
        if started:
            last = started.bit_length() - 1
            carry_lo, carry_hi = (lo[31] - excl_lo[last]) & FIELDS, (hi[31] - excl_hi[last]) & FIELDS
        else:
            carry_lo, carry_hi = (carry_lo + lo[31]) & FIELDS, (carry_hi + hi[31]) & FIELDS


def hacc_strip(mem, addr, length, row_bytes, spp, warps=4):
    """slzw_hacc_kernel for one strip: its rows in `warps` contiguous groups."""
    rows = (length + row_bytes - 1) // row_bytes
    for w in range(warps):
        r0, r1 = rows * w // warps, rows * (w + 1) // warps
        if r1 > r0:
            a = r0 * row_bytes
            e = min(length, r1 * row_bytes)
            accumulate_rows(mem, addr + a, e - a, row_bytes, spp)
