"""Pin oracle/philox.py to the Random123 known-answer vectors (philox4x32_10, kat_vectors)."""
import numpy as np

from oracle import philox as px

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        got = px.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_vectorised_equals_scalar_and_addressing():
    cell = np.array([0, 1, 2 ** 32 + 5, 2 ** 40], dtype=np.uint64)
    w = px.words(seed=0x1234567890ABCDEF, it=7, purpose=px.PUR_E, cell=cell, sub=3)
    for i, c in enumerate(cell):
        c = int(c)
        s = px.philox4x32_10(c & 0xFFFFFFFF, c >> 32, 3, (7 << 8) | px.PUR_E, 0x90ABCDEF, 0x12345678)
        assert [int(x[i]) for x in w] == [int(x) for x in s]


def test_u01_open_interval():
    w = np.array([0, 1, 0xFFFFFFFF], dtype=np.uint32)
    for bits in (32, 24):
        u = px.u01(w, bits)
        assert (u > 0).all() and (u < 1).all()
    assert px.u01(np.uint32(0)) == 2.0 ** -33
