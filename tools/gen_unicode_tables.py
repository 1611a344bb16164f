"""Prints the Unicode tables of csrc/scan_core.cuh from this interpreter's unicodedata and checks
that they describe exactly what `re`'s \\d / \\s and float() accept (Python 3.12 = Unicode 15.0)."""
import re
import unicodedata


def nd_runs():
    starts, cp = [], 0x80
    while cp < 0x110000:
        if unicodedata.decimal(chr(cp), -1) != -1:
            s = cp
            while cp < 0x110000 and unicodedata.decimal(chr(cp), -1) != -1:
                cp += 1
            assert (cp - s) % 10 == 0
            assert all(unicodedata.decimal(chr(s + i)) == i % 10 for i in range(cp - s))
            starts += list(range(s, cp, 10))
        else:
            cp += 1
    return starts


if __name__ == "__main__":
    starts = nd_runs()
    digits = {s + i for s in starts for i in range(10)} | set(range(0x30, 0x3A))
    assert all((re.match(r"\d", chr(c)) is not None) == (c in digits) for c in range(0x110000))
    spaces = [c for c in range(0x110000) if chr(c).isspace()]
    assert all((re.match(r"\s", chr(c)) is not None) == (c in spaces) for c in range(0x110000))
    print("unicode", unicodedata.unidata_version, "runs", len(starts))
    print("{" + ", ".join("0x%04X" % s for s in starts) + "}")
    print("spaces", [hex(c) for c in spaces])
