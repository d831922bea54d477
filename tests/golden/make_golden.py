"""Regenerate the committed fixtures under tests/golden/ from the reference's bundled data.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):      python tests/golden/make_golden.py

  COSMIC_v3.3.1_SBS_GRCh37.csv   verbatim data fixture, inst/extdata/ of the reference
                                 (96 x 79 signatures, columns sum to 1)
  example_data.npz               inst/extdata/example_data.rds decoded: M int 96 x 64,
                                 P 96 x 4 (= COSMIC SBS58, SBS40, SBS26, SBS2),
                                 E int 4 x 64, probs 4 x 64   

R is not installed here, so the RDS (gzip + R serialization v3, XDR) is decoded with
the minimal reader below: lists, integer / real / character vectors and attribute
pairlists are all the file contains.
"""
import gzip
import os
import shutil
import struct

import numpy as np

REF = "/root/reference/inst/extdata"
HERE = os.path.dirname(os.path.abspath(__file__))


class XDR:
    def __init__(self, b):
        self.b, self.o = b, 0

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64(self, n):
        v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype(np.float64)
        self.o += 8 * n
        return v

    def i32s(self, n):
        v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype(np.int64)
        self.o += 4 * n
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


def read_item(x, symtab):
    flags = x.i32()
    ty = flags & 0xFF
    has_attr = bool(flags & 0x200)
    has_tag = bool(flags & 0x400)
    if ty == 254:                      # NILVALUE
        return None
    if ty == 255:                      # REFSXP
        return symtab[(flags >> 8) - 1]
    if ty == 1:                        # SYMSXP
        name = read_item(x, symtab)
        symtab.append(name)
        return name
    if ty == 2:                        # LISTSXP (pairlist): attr? tag? car cdr
        out = {}
        while True:
            attr = read_item(x, symtab) if has_attr else None  # noqa: F841
            tag = read_item(x, symtab) if has_tag else None
            car = read_item(x, symtab)
            out[tag] = car
            flags = x.i32()
            ty = flags & 0xFF
            has_attr = bool(flags & 0x200)
            has_tag = bool(flags & 0x400)
            if ty == 254:
                return out
            assert ty == 2, ty
    if ty == 9:                        # CHARSXP
        n = x.i32()
        return None if n == -1 else x.raw(n).decode()
    if ty in (10, 13):                 # LGLSXP / INTSXP
        val = x.i32s(x.i32())
    elif ty == 14:                     # REALSXP
        val = x.f64(x.i32())
    elif ty == 16:                     # STRSXP
        val = [read_item(x, symtab) for _ in range(x.i32())]
    elif ty == 19:                     # VECSXP
        val = [read_item(x, symtab) for _ in range(x.i32())]
    else:
        raise NotImplementedError(f"SEXP type {ty}")
    attrs = read_item(x, symtab) if has_attr else {}
    if isinstance(val, np.ndarray) and attrs and "dim" in attrs:
        val = val.reshape([int(d) for d in attrs["dim"]], order="F")
    if isinstance(val, list) and attrs and "names" in attrs:
        val = dict(zip(attrs["names"], val))
    return val


def read_rds(path):
    b = gzip.open(path).read()
    assert b[:2] == b"X\n"
    x = XDR(b)
    x.o = 2
    version = x.i32(); x.i32(); x.i32()
    if version == 3:
        x.raw(x.i32())                 # native encoding
    return read_item(x, [])


def main():
    shutil.copyfile(os.path.join(REF, "COSMIC_v3.3.1_SBS_GRCh37.csv"),
                    os.path.join(HERE, "COSMIC_v3.3.1_SBS_GRCh37.csv"))
    d = read_rds(os.path.join(REF, "example_data.rds"))
    out = {k: np.asarray(d[k]) for k in ("M", "P", "E", "probs") if k in d}
    np.savez_compressed(os.path.join(HERE, "example_data.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(v.mean()))


if __name__ == "__main__":
    main()
