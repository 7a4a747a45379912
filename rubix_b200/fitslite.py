"""Minimal FITS writer / reader (numpy only; the image has no astropy): a header-only primary HDU followed by
IMAGE extensions, which is all rubix/core/fits.py:13-101 writes.  FITS standard 4.0: 80-character cards,
2880-byte blocks, big-endian data, BITPIX -32 / -64 / 16 / 32 / 64."""

from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

BLOCK = 2880
_BITPIX = {np.dtype(">f4"): -32, np.dtype(">f8"): -64, np.dtype(">i2"): 16, np.dtype(">i4"): 32, np.dtype(">i8"): 64,
           np.dtype("u1"): 8}
_DTYPE = {v: k for k, v in _BITPIX.items()}


def _card(key: str, value, comment: str = "") -> bytes:
    key = key.upper()
    if len(key) > 8:
        raise ValueError(f"FITS keyword {key!r} longer than 8 characters")
    if isinstance(value, (bool, np.bool_)):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        r = repr(float(value)).upper()
        v = f"{r:>20}"
    else:
        t = str(value).replace("'", "''")
        v = f"'{t:<8}'"
        v = f"{v:<20}"
    s = f"{key:<8}= {v}"
    if comment:
        s += f" / {comment}"
    if len(s) > 80:
        raise ValueError(f"FITS card too long: {s!r}")
    return s.ljust(80).encode("ascii")


def _header_bytes(cards: List[Tuple[str, object]]) -> bytes:
    raw = b"".join(_card(k, v) for k, v in cards) + b"END".ljust(80)
    return raw + b" " * ((-len(raw)) % BLOCK)


def write_fits(path: str, primary_header: Dict[str, object], images: List[Tuple[np.ndarray, Dict[str, object]]]) -> None:
    """Header-only primary HDU + one IMAGE extension per (array, header)."""
    out = [_header_bytes([("SIMPLE", True), ("BITPIX", 8), ("NAXIS", 0), ("EXTEND", True)]
                         + [(k, v) for k, v in primary_header.items() if k.upper() != "SIMPLE"])]
    for arr, hdr in images:
        a = np.asarray(arr)
        be = a.astype(a.dtype.newbyteorder(">"), copy=False) if a.dtype.kind in "fi" else a
        if be.dtype not in _BITPIX:
            raise ValueError(f"unsupported FITS dtype {a.dtype}")
        cards = [("XTENSION", "IMAGE"), ("BITPIX", _BITPIX[be.dtype]), ("NAXIS", a.ndim)]
        cards += [(f"NAXIS{i + 1}", n) for i, n in enumerate(reversed(a.shape))]   # NAXIS1 is the fastest axis
        cards += [("PCOUNT", 0), ("GCOUNT", 1)] + list(hdr.items())
        data = np.ascontiguousarray(be).tobytes()
        out += [_header_bytes(cards), data + b"\0" * ((-len(data)) % BLOCK)]
    with open(path, "wb") as f:
        for b in out:
            f.write(b)


def _parse_value(s: str):
    s = s.split(" / ")[0].strip() if not s.strip().startswith("'") else s.strip()
    if s.startswith("'"):
        end = s.rfind("'")
        return s[1:end].replace("''", "'").rstrip()
    if s in ("T", "F"):
        return s == "T"
    try:
        return int(s)
    except ValueError:
        return float(s)


def read_fits(path: str) -> List[Tuple[Dict[str, object], Optional[np.ndarray]]]:
    """[(header, data or None)] for every HDU written by :func:`write_fits` (or any simple image FITS)."""
    buf = open(path, "rb").read()
    pos, hdus = 0, []
    while pos < len(buf):
        hdr: Dict[str, object] = {}
        done = False
        while not done:
            block = buf[pos:pos + BLOCK]
            if len(block) < BLOCK:
                raise ValueError("truncated FITS header")
            pos += BLOCK
            for i in range(0, BLOCK, 80):
                card = block[i:i + 80].decode("ascii")
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if card[8:10] == "= ":
                    val = card[10:]
                    if val.strip().startswith("'"):
                        q = val.find("'", val.find("'") + 1)
                        while q + 1 < len(val) and val[q + 1] == "'":
                            q = val.find("'", q + 2)
                        val = val[:q + 1]
                    hdr[key] = _parse_value(val)
        naxis = int(hdr.get("NAXIS", 0))
        shape = tuple(int(hdr[f"NAXIS{i}"]) for i in range(naxis, 0, -1))
        data = None
        if naxis and int(np.prod(shape)):
            dt = _DTYPE[int(hdr["BITPIX"])]
            nbytes = int(np.prod(shape)) * dt.itemsize
            data = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape)), offset=pos).reshape(shape)
            data = data.astype(dt.newbyteorder("="))
            pos += nbytes + ((-nbytes) % BLOCK)
        hdus.append((hdr, data))
    return hdus
