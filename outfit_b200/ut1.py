"""Host-side reader of JPL's Earth-orientation file `latest_eop2.long` -> UT1 epochs for `pvobs`.

The reference gets UT1 from hifitime (un-vendored): `Ut1Provider::download_from_jpl("latest_eop2.long")`
(examples/run_full_iod.rs:111, tests/test_gauss_iod.rs:76) and, per observation,
`tmjd.to_ut1(ut1_provider).to_mjd_tai_days()` (observer_extension.rs:191-192).  This module restates
that chain from hifitime 4's published behaviour so a downloaded EOP2 file can feed the batch's
`mjd_ut1` column without Rust:

  * the file is a namelist: everything before the line ` EOP2=` is header, the data end at ` $END`;
    each data line is comma separated, column 0 = MJD (of a TAI-scale epoch), column 3 = TAI - UT1 in
    milliseconds (columns 1-2 are the polar motion, the rest are sigmas and correlations);
  * `ut1_offset(epoch)`: the table is scanned from the END and the first entry whose epoch is strictly
    EARLIER than the query wins -- a step function, no interpolation; no such entry -> offset 0;
  * `to_ut1` = (epoch in TAI) - offset; `to_mjd_tai_days` = that instant as MJD days.
  * hifitime keeps integer nanoseconds: the MJD -> Duration conversion and the millisecond offsets are
    rounded to 1 ns here as well.

Parity with the crate is UNPINNED (no hifitime source, no EOP file and no DE440 in this image; the
reference's tests that would pin it need both downloads): DESIGN.md 8.  The arithmetic below is exact
integer nanoseconds, so the only freedom left is hifitime's rounding of the f64 inputs.
"""
import bisect

import numpy as np

NS_PER_DAY = 86_400_000_000_000
TT_MINUS_TAI_NS = 32_184_000_000  # 32.184 s


def _days_to_ns(days):
    return int(round(float(days) * NS_PER_DAY))


class Ut1Table:
    """The (epoch, TAI - UT1) step table of a JPL EOP2 file."""

    def __init__(self, mjd_tai, tai_minus_ut1_ms):
        mjd = np.asarray(mjd_tai, dtype=np.float64)
        ms = np.asarray(tai_minus_ut1_ms, dtype=np.float64)
        if mjd.shape != ms.shape or mjd.ndim != 1:
            raise ValueError("mjd_tai and tai_minus_ut1_ms must be 1-D arrays of equal length")
        self.epoch_ns = [_days_to_ns(d) for d in mjd]          # file order (hifitime scans it reversed)
        self.offset_ns = [int(round(float(v) * 1e6)) for v in ms]
        self._sorted = all(a <= b for a, b in zip(self.epoch_ns, self.epoch_ns[1:]))

    @classmethod
    def from_eop2_text(cls, text):
        mjd, ms = [], []
        ignore = True
        for line in text.splitlines():
            if line == " EOP2=":
                ignore = False
                continue
            if line == " $END":
                break
            if ignore:
                continue
            cols = line.split(",")
            if len(cols) < 4:
                raise ValueError(f"EOP2 data line with {len(cols)} columns: {line!r}")
            mjd.append(float(cols[0].strip()))
            ms.append(float(cols[3].strip()))
        if ignore:
            raise ValueError("no ' EOP2=' marker: not a JPL EOP2 file")
        return cls(mjd, ms)

    @classmethod
    def from_file(cls, path):
        with open(path, "r") as f:
            return cls.from_eop2_text(f.read())

    def __len__(self):
        return len(self.epoch_ns)

    def offset_ns_at(self, tai_ns):
        """TAI - UT1 (ns) hifitime's Epoch::ut1_offset returns for a TAI instant (ns since MJD 0)."""
        if self._sorted:
            i = bisect.bisect_left(self.epoch_ns, tai_ns)  # entries [0, i) are strictly earlier
            return self.offset_ns[i - 1] if i > 0 else 0
        for e, o in zip(reversed(self.epoch_ns), reversed(self.offset_ns)):
            if tai_ns > e:
                return o
        return 0

    def mjd_ut1(self, mjd_tt):
        """`Epoch::from_mjd_in_time_scale(mjd_tt, TT).to_ut1(self).to_mjd_tai_days()` for an array of MJD(TT)."""
        mjd_tt = np.atleast_1d(np.asarray(mjd_tt, dtype=np.float64))
        out = np.empty_like(mjd_tt)
        for i, d in enumerate(mjd_tt):
            tai = _days_to_ns(d) - TT_MINUS_TAI_NS
            ut1 = tai - self.offset_ns_at(tai)
            whole, frac = divmod(ut1, NS_PER_DAY)
            out[i] = whole + frac / NS_PER_DAY
        return out

    def dut1_seconds(self, mjd_tt, tai_minus_utc_s):
        """UT1 - UTC (s) at the given epochs, for callers that carry UTC (tai_minus_utc_s = leap seconds)."""
        mjd_tt = np.atleast_1d(np.asarray(mjd_tt, dtype=np.float64))
        lead = np.broadcast_to(np.asarray(tai_minus_utc_s, dtype=np.float64), mjd_tt.shape)
        return np.array([lead[i] - self.offset_ns_at(_days_to_ns(d) - TT_MINUS_TAI_NS) * 1e-9 for i, d in enumerate(mjd_tt)])
