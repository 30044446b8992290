"""Host-side reader for JPL DE binary ephemeris files (little-endian `linux_p1550p2650.440` family) ->
the Chebyshev table of `outfit_b200_load_ephemeris`.  SURVEY 8f-1: lets a real DE440 be consumed
without the Rust side.

Layout followed (reference `HorizonData::read_horizon_file`,
/root/reference/src/jpl_ephem/horizon/horizon_data.rs:239-251, 270-292, 336-368, 598-684): header
record = TTL (14*3 x 6 chars), CNAM (400 x 6 chars), SS[3] f64 (start JD, end JD, days per record),
NCON i32, AU f64, EMRAT f64 at byte 2688, IPT[12][3] u32 + NUMDE u32 + LPT[3] u32 at byte 2696,
IPT[13], IPT[14] after the constant names when NCON > 400 (DE440 and later); record size =
(4 + sum 2 * n_coeff * n_sub * dim) * 4 bytes; data records start at byte 2 * recsize, each
[jd_start, jd_end, coefficients ...] with IPT offsets 1-based from the start of the record.
The three bodies the observer position needs: IPT rows 2 (Earth-Moon barycentre), 9 (Moon,
geocentric) and 10 (Sun), horizon_ids.rs:35-45.

No real DE file exists in this image: the reader is exercised on files written by `write_de_binary`
from the synthetic table (tests/test_de_reader.py), i.e. the LAYOUT is pinned by the reference's
parser, the numbers are synthetic.
"""
import struct

import numpy as np

_DIM = [3] * 11 + [2, 3, 3, 1]  # nutation (row 11) has 2 components, TT-TDB (row 14) one


def _recsize(ipt):
    return (4 + sum(2 * int(ipt[i][1]) * int(ipt[i][2]) * _DIM[i] for i in range(15))) * 4


def read_de_binary(path):
    """-> dict(cheb [n_blocks, ncoeff] f64, jd_start, block_days, ipt uint32[3,3] (0-based offset, n_coeff,
    n_sub for EMB, Moon, Sun), emrat, numde, jd_end, ipt_full)."""
    with open(path, "rb") as f:
        head = f.read(1 << 12)
        ss = struct.unpack_from("<3d", head, 2652)
        ncon = struct.unpack_from("<i", head, 2676)[0]
        emrat = struct.unpack_from("<d", head, 2688)[0]
        raw = struct.unpack_from("<40I", head, 2696)
        ipt = [[0, 0, 0] for _ in range(15)]
        for i in range(36):
            ipt[i // 3][i % 3] = raw[i]
        numde = raw[36]
        ipt[12] = list(raw[37:40])
        if numde >= 440 and ncon > 400:
            f.seek(2856 + (ncon - 400) * 6)
            extra = struct.unpack("<6I", f.read(24))
            ipt[13], ipt[14] = list(extra[:3]), list(extra[3:])
        recsize = _recsize(ipt)
        ncoeff = recsize // 8
        f.seek(2 * recsize)
        data = np.frombuffer(f.read(), dtype="<f8")
    n_blocks = data.size // ncoeff
    cheb = np.ascontiguousarray(data[:n_blocks * ncoeff].reshape(n_blocks, ncoeff))
    if n_blocks == 0:
        raise ValueError("no data records")
    if abs(cheb[0, 0] - ss[0]) > 1e-6 or abs((cheb[0, 1] - cheb[0, 0]) - ss[2]) > 1e-6:
        raise ValueError("first data record does not start at SS[0] / span SS[2] days: not a little-endian DE file?")
    sel = np.array([[ipt[b][0] - 1, ipt[b][1], ipt[b][2]] for b in (2, 9, 10)], dtype=np.uint32)
    return {"cheb": cheb, "jd_start": float(ss[0]), "jd_end": float(ss[1]), "block_days": float(ss[2]), "ipt": sel,
            "emrat": float(emrat), "numde": int(numde), "ipt_full": ipt}


def write_de_binary(path, table, numde=440):
    """Write the synthetic Chebyshev table (outfit_b200.synth.make_ephemeris_table) as a DE-layout file
    holding only the EMB, Moon and Sun rows (the other IPT rows are empty).  Test helper."""
    cheb = np.asarray(table["cheb"], dtype=np.float64)
    n_blocks, stride = cheb.shape
    tip = np.asarray(table["ipt"], dtype=np.int64)
    ipt = [[0, 0, 0] for _ in range(15)]
    # pack the three bodies back to back after the two JD doubles (1-based offsets)
    off = 3
    order = {2: 0, 9: 1, 10: 2}
    for row, b in order.items():
        ipt[row] = [off, int(tip[b][1]), int(tip[b][2])]
        off += 3 * int(tip[b][1]) * int(tip[b][2])
    recsize = _recsize(ipt)
    ncoeff = recsize // 8
    assert off - 1 == ncoeff, (off, ncoeff)
    jd0, days = float(table["jd_start"]), float(table["block_days"])
    head = bytearray(max(2 * recsize, 1 << 12))
    head[0:6] = b"SYNTH "
    struct.pack_into("<3d", head, 2652, jd0, jd0 + days * n_blocks, days)
    struct.pack_into("<i", head, 2676, 0)
    struct.pack_into("<d", head, 2680, 149597870.7)
    struct.pack_into("<d", head, 2688, float(table["emrat"]))
    flat = [v for row in ipt[:12] for v in row]
    struct.pack_into("<36I", head, 2696, *flat)
    struct.pack_into("<I", head, 2696 + 144, numde)
    struct.pack_into("<3I", head, 2696 + 148, *ipt[12])
    rec = np.zeros((n_blocks, ncoeff))
    rec[:, 0] = jd0 + days * np.arange(n_blocks)
    rec[:, 1] = rec[:, 0] + days
    for row, b in order.items():
        n = 3 * int(tip[b][1]) * int(tip[b][2])
        rec[:, ipt[row][0] - 1: ipt[row][0] - 1 + n] = cheb[:, int(tip[b][0]): int(tip[b][0]) + n]
    with open(path, "wb") as f:
        f.write(bytes(head[:2 * recsize]))
        f.write(rec.astype("<f8").tobytes())
