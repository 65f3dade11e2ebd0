"""What the design relies on in the generated code, checked on the built objects with cuobjdump
(no GPU needed): the default encoder kernel keeps its dictionaries in tensor memory and shared
memory (LDTM / STTM / LDSM), finds a key with one warp reduction (REDUX), the per-byte loops of the codec kernels do not spill, and the
shared-memory lookup step carries no divergence guard (profiles/r01_encode_notes.md)."""
import os
import re
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

CSRC = os.path.join(ROOT, "lzw_b200", "csrc")


def _functions(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(os.path.join(CSRC, obj)) or not os.path.exists(exe):
        pytest.skip("built objects or cuobjdump not available")
    out = subprocess.run([exe, "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    return funcs


def _ops(lines):
    ops = []
    for line in lines:
        body = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
        toks = [t for t in body.split() if not t.startswith("@")]
        if toks:
            ops.append(toks[0])
    return ops


def test_default_encoder_uses_tensor_memory_and_matrix_loads():
    funcs = _functions("encode_kernels.o")
    # slzw_encode_kernel<96, 12, 16, 2, FIXED>: the one encoder kernel of the library
    default = {n: l for n, l in funcs.items() if "slzw_encode_kernelILi96ELi12ELi16ELi2ELb" in n}
    assert len(default) == 2
    # + the latency variant of the fixed flavour; no experiment kernels in the product build
    assert len([n for n in funcs if "slzw_encode" in n]) == 3
    for name, lines in default.items():
        ops = _ops(lines)
        assert any(o.startswith("LDTM") for o in ops), name          # tcgen05.ld
        assert any(o.startswith("STTM") for o in ops), name          # tcgen05.st
        assert any(o.startswith("LDSM") for o in ops), name          # ldmatrix bucket loads
        assert any("REDUX" in o for o in ops), name                  # hit detection: one warp min-reduction
        # no spills in the per-byte loops: at most one loop-invariant address parked once at kernel
        # start (one STL) and fetched once per tile in the code-packing tail (one LDL)
        assert sum(o.startswith(("LDL", "STL")) for o in ops) <= 2, name
        # the instruction after an LDSM-based lookup reaches its ballot without a divergence guard
        idx = [i for i, o in enumerate(ops) if o.startswith("LDSM")]
        guarded = sum(1 for i in idx if any(o == "BRA.DIV" for o in ops[i:i + 12]))
        assert guarded == 0, (name, guarded)


def test_codec_kernels_do_not_spill():
    for obj, key in (("decode_kernels.o", "slzw_decode_fast_kernelILi12ELi20E"), ("sched_kernels.o", "")):
        for name, lines in _functions(obj).items():
            if key in name:
                assert not any(o.startswith(("LDL", "STL")) for o in _ops(lines)), name


def test_lookup_loops_stay_within_their_instruction_budget():
    """The encoder is bound by instruction issue; the four-lookup loop bodies of the MODE 0 loops
    (the code nearly every input byte runs through) were 169 / 161 SASS instructions (tensor / shared
    memory) at 71.3 ms per launch and are 156 / 141 at 65.5 ms (profiles/r02_encode_notes.md).  A
    change that lets the compiler put instructions back fails here before it costs GPU time."""
    funcs = _functions("encode_kernels.o")
    name = next(n for n in funcs if "slzw_encode_kernelILi96ELi12ELi16ELi2ELb0ELb0" in n)
    ins = []
    for line in funcs[name]:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    index = {a: i for i, (a, _) in enumerate(ins)}
    best = {"LDTM": None, "LDSM": None}
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
        if not m or "BRA.DIV" in text:
            continue
        t = int(m.group(1), 16)
        if t >= a or t not in index:
            continue
        body = [x for _, x in ins[index[t]:i + 1]]
        for kind in best:
            # four lookups + their four overflow reloads, one insert per lookup
            loads = sum(1 for x in body if re.search(r"\b%s" % kind, x))
            inserts = sum(1 for x in body if re.search(r"\bSTTM\b" if kind == "LDTM" else r"\bSTS\b", x))
            if loads == 8 and (inserts == 4 if kind == "LDTM" else inserts == 8):
                if best[kind] is None or len(body) < best[kind]:
                    best[kind] = len(body)
    assert best["LDTM"] is not None and best["LDSM"] is not None, best
    assert best["LDTM"] <= 160, best
    assert best["LDSM"] <= 146, best
