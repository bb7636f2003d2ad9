"""Minimal reader for R's RDX3 (XDR) serialisation as used by the reference's data/*.rda files (plain data frames).
Used only by tools/make_golden.py to turn reference data into small committed fixtures; nothing at run time reads R files."""
import bz2
import gzip
import struct

import numpy as np


class _R:
    def __init__(self, b):
        self.b, self.p = b, 0

    def i4(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v


def _item(r, refs):
    flags = r.i4()
    t = flags & 0xFF
    has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
    if t == 254:      # NILVALUE
        return None
    if t == 255:      # REFSXP
        idx = flags >> 8
        if idx == 0:
            idx = r.i4()
        return refs[idx - 1]
    if t == 1:        # SYMSXP
        s = _item(r, refs)
        refs.append(s)
        return s
    if t == 9:        # CHARSXP
        n = r.i4()
        return None if n == -1 else r.raw(n).decode("utf-8", "replace")
    if t in (2, 6):   # LISTSXP / LANGSXP (pairlist)
        out = {}
        attr = _item(r, refs) if has_attr else None
        tag = _item(r, refs) if has_tag else None
        car = _item(r, refs)
        out[tag if tag is not None else len(out)] = car
        cdr = _item(r, refs)
        if isinstance(cdr, dict):
            out.update(cdr)
        return out
    if t == 10:       # LGLSXP
        n = r.i4()
        v = np.frombuffer(r.raw(4 * n), dtype=">i4").astype(np.int32)
    elif t == 13:     # INTSXP
        n = r.i4()
        v = np.frombuffer(r.raw(4 * n), dtype=">i4").astype(np.int32)
    elif t == 14:     # REALSXP
        n = r.i4()
        v = np.frombuffer(r.raw(8 * n), dtype=">f8").astype(np.float64)
    elif t == 16:     # STRSXP
        n = r.i4()
        v = [_item(r, refs) for _ in range(n)]
    elif t == 19:     # VECSXP
        n = r.i4()
        v = [_item(r, refs) for _ in range(n)]
    elif t == 238:    # ALTREP
        info = _item(r, refs); state = _item(r, refs); _ = _item(r, refs)
        cls = list(info.values())[0] if isinstance(info, dict) else info
        if cls == "compact_intseq":
            n, start, step = (int(x) for x in state)
            v = (start + step * np.arange(n)).astype(np.int32)
        elif cls == "wrap_real" or cls == "wrap_integer" or cls == "wrap_string":
            v = state[0] if isinstance(state, list) else list(state.values())[0]
        else:
            raise NotImplementedError(f"ALTREP class {cls}")
        return v
    else:
        raise NotImplementedError(f"SEXP type {t} at {r.p}")
    if has_attr:
        attr = _item(r, refs)
        return {"value": v, "attr": attr}
    return v


def read_rda(path):
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    assert raw[:5] == b"RDX3\n", raw[:8]
    r = _R(raw)
    r.p = 5
    assert r.raw(2) == b"X\n"
    r.i4(); r.i4(); r.i4()              # version, writer version, min reader version
    n = r.i4(); r.raw(n)                 # native encoding
    return _item(r, [])


def data_frame(obj):
    """{column: array} from a serialised data.frame (first object of an .rda pairlist)."""
    df = list(obj.values())[0] if isinstance(obj, dict) and "value" not in obj else obj
    names = df["attr"]["names"]
    names = names["value"] if isinstance(names, dict) else names
    cols = {}
    for nm, c in zip(names, df["value"]):
        cols[nm] = c["value"] if isinstance(c, dict) else c
    return cols
