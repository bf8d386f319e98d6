"""Generates tests/golden/updown_headers.npz: the header blocks of SOS_Up.txt / SOS_Down.txt exactly as the reference's
SOS_OUTPUT_HEADER and SOS_OUTPUT_HEADER_POLAR_DIAG write them (SOS_TRPHI.F:1570-1796).

The translated reference library drops message WRITEs, so the headers are taken from the Fortran source itself: this
script reads the WRITE statements of the two routines where they lie under /root/reference (nothing is copied into the
repository but the resulting header text, i.e. the output FORMAT the drop-in writer has to reproduce), applies Fortran's
Aw / Fw.d output rules on BYTES (the source is UTF-8: the degree sign counts two characters, as gfortran counts it) and
stores the result for a few argument sets.  Run here (needs /root/reference); the GPU box only reads the .npz."""
import os
import re
import sys

import numpy as np

REF = "/root/reference/src/SOS_TRPHI.F"
HERE = os.path.dirname(os.path.abspath(__file__))


def routine_statements(name):
    """Logical statements (continuations joined, comments dropped) of SUBROUTINE `name`, as bytes."""
    raw = open(REF, "rb").read().split(b"\n")
    out, cur, inside = [], None, False
    for ln in raw:
        ln = ln.rstrip(b"\r")
        if not ln.strip() or ln[:1] in (b"C", b"c", b"*", b"!"):
            continue
        ln = ln.replace(b"\t", b"      ", 1) if ln.startswith(b"\t") else ln
        body = ln[6:] if len(ln) > 6 else b""
        cont = len(ln) > 5 and ln[5:6] not in (b" ", b"0") and not ln[:5].strip()
        if cont and cur is not None:
            cur += body
            continue
        if cur is not None:
            out.append(cur)
        cur = ln[:5] + b" " + body          # keep the label field
    if cur is not None:
        out.append(cur)
    res = []
    for st in out:
        txt = st[6:].strip()
        if re.match(rb"SUBROUTINE\s+" + name.encode() + rb"\b", txt):
            inside = True
            continue
        if inside:
            if re.match(rb"END\b", txt):
                break
            res.append((st[:5].strip(), txt))
    return res


def literals(expr):
    """'a' // 'b' ... -> concatenated bytes, and the remaining (non-character) items."""
    s, items, i = b"", [], 0
    parts = re.findall(rb"'((?:[^']|'')*)'|([^,/' ][^,']*)", expr)
    for lit, other in parts:
        if lit or (not other.strip()):
            s += lit.replace(b"''", b"'")
        elif other.strip():
            items.append(other.strip())
    return s, items


def fmt_a(s, w):
    return s[:w] if len(s) >= w else b" " * (w - len(s)) + s


def fmt_f(x, w, d):
    t = ("%.*f" % (d, x)).encode()
    return b" " * (w - len(t)) + t if len(t) <= w else b"*" * w


def render(name, values):
    lines = []
    stmts = routine_statements(name)
    formats = {lab: txt for lab, txt in stmts if txt.upper().startswith(b"FORMAT")}
    updown = values["UPDOWN"]
    stack = []                                   # IF (UPDOWN.EQ.1) THEN / ELSE / ENDIF
    for lab, txt in stmts:
        up = txt.upper()
        if up.startswith(b"IF") and b"THEN" in up:
            stack.append(updown == 1)
            continue
        if up == b"ELSE":
            stack[-1] = not stack[-1]
            continue
        if up in (b"ENDIF", b"END IF"):
            stack.pop()
            continue
        if not up.startswith(b"WRITE") or (stack and not all(stack)):
            continue
        m = re.match(rb"WRITE\s*\(\s*FICID\s*,\s*('\(A(\d+)\)'|(\d+))\s*\)\s*(.*)$", txt, re.S)
        assert m, txt
        s, items = literals(m.group(4))
        if m.group(2):
            lines.append(fmt_a(s, int(m.group(2))))
        else:
            f = formats[m.group(3)]
            mm = re.match(rb"FORMAT\s*\(\s*A(\d+)\s*,\s*1X\s*,\s*F(\d+)\.(\d+)\s*\)", f)
            assert mm and len(items) == 1, (f, items)
            lines.append(fmt_a(s, int(mm.group(1))) + b" " + fmt_f(values[items[0].decode().strip()], int(mm.group(2)), int(mm.group(3))))
    return b"\n".join(lines) + b"\n"


def main():
    if not os.path.exists(REF):
        sys.exit("needs /root/reference")
    out = {}
    for tag, phi1, zalt_up, zalt_dn in (("toa", 0.0, 120.0, 0.0), ("z3_phi20", 20.0, 3.0, 3.0)):
        for ud, z in ((1, zalt_up), (2, zalt_dn)):
            out["view1_%s_%d" % (tag, ud)] = np.frombuffer(render("SOS_OUTPUT_HEADER", dict(UPDOWN=ud, PHI1=phi1, PHI2=phi1 + 180.0, ZALT=z)), dtype=np.uint8)
            out["view2_%s_%d" % (tag, ud)] = np.frombuffer(render("SOS_OUTPUT_HEADER_POLAR_DIAG", dict(UPDOWN=ud, ZALT=z)), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden", "updown_headers.npz"), **out)
    print(out["view1_toa_1"].tobytes().decode())
    print(out["view2_toa_2"].tobytes().decode())


if __name__ == "__main__":
    main()
