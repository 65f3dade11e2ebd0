#!/usr/bin/env python
"""Lists the loops (backward branches) of the encode kernels in an object file with their size in
instructions and the bucket loads / reductions / inserts they contain -- a quick way to compare the
per-byte instruction count of two builds without a GPU.

    python tools/sass_loops.py lzw_b200/csrc/encode_kernels.o [kernel-name-substring]
"""
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "slzw_encode_kernel"
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if want not in name:
            continue
        short = re.search(r"slzw_encode_kernel(\w+?)EvNS", name)
        print(f"== {short.group(1) if short else name}: {len(ins)} instructions")
        addr = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, text) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
            if not m or "BRA.DIV" in text:
                continue
            t = int(m.group(1), 16)
            if t >= a or t not in addr:
                continue
            body = [x for _, x in ins[addr[t]:i + 1]]
            cnt = lambda k: sum(1 for x in body if re.search(k, x))
            loads = cnt(r"\bLDTM\b") + cnt(r"\bLDSM")
            if loads < 2:
                continue
            print(f"  loop {t:#06x}..{a:#06x}: {len(body):4d} instr, LDTM {cnt(r'LDTM')}, LDSM {cnt(r'LDSM')}, "
                  f"REDUX {cnt(r'REDUX')}, STTM {cnt(r'STTM')}, STS {cnt(r'STS')}, VOTE {cnt(r'VOTE')}, R2UR {cnt(r'R2UR')}")


if __name__ == "__main__":
    main()
