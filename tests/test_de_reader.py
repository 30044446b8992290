"""DE binary reader: layout of the reference's parser, synthetic numbers."""
import numpy as np
import pytest


def test_de_binary_roundtrip_and_earth_positions(tmp_path, oracle):
    from outfit_b200 import de_reader, synth
    table = synth.make_ephemeris_table(n_blocks=12)
    path = str(tmp_path / "synth.440")
    de_reader.write_de_binary(path, table)
    got = de_reader.read_de_binary(path)
    assert got["numde"] == 440 and got["block_days"] == table["block_days"] and got["jd_start"] == table["jd_start"]
    assert got["emrat"] == table["emrat"] and got["cheb"].shape[0] == 12
    assert [list(r[1:]) for r in got["ipt"]] == [list(r[1:]) for r in np.asarray(table["ipt"])]
    # header fields at the byte offsets the reference reads (horizon_data.rs:620-650)
    raw = open(path, "rb").read()
    assert np.frombuffer(raw[2652:2676], "<f8")[2] == table["block_days"] and np.frombuffer(raw[2688:2696], "<f8")[0] == table["emrat"]
    # the table read from the file evaluates to the same Earth positions, bit for bit
    a = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    b = oracle.make_ephem_table(got["cheb"], got["jd_start"], got["block_days"], got["ipt"], got["emrat"])
    import ctypes as C
    for t in np.linspace(58000.3, 58000.0 + 32 * 12 - 0.7, 57):
        pa, pb, v = oracle.D3(), oracle.D3(), oracle.D3()
        assert oracle.lib().oo_earth_ephemeris(C.byref(a), float(t), 0, pa, v) == 0
        assert oracle.lib().oo_earth_ephemeris(C.byref(b), float(t), 0, pb, v) == 0
        assert list(pa) == list(pb)
    with pytest.raises(ValueError):
        bad = bytearray(raw); bad[2652:2660] = np.float64(1.0).tobytes()
        p2 = str(tmp_path / "bad.440"); open(p2, "wb").write(bytes(bad)); de_reader.read_de_binary(p2)


@pytest.mark.gpu
def test_gpu_accepts_table_read_from_de_file(tmp_path, oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import IODParams, OutfitB200, de_reader, synth
    table = synth.make_ephemeris_table()
    path = str(tmp_path / "synth.440")
    de_reader.write_de_binary(path, table)
    batch = synth.make_trajectories(64, 10, seed=151, table=table, max_triplets=8, n_noise=1)
    p = IODParams.builder(n_noise_realizations=0, max_triplets=8)
    a, b = OutfitB200(0), OutfitB200(0)
    a.load_ephemeris(table)
    b.load_ephemeris(de_reader.read_de_binary(path))
    assert a.fit_full_iod(batch, p, use_body_fixed=True).tobytes() == b.fit_full_iod(batch, p, use_body_fixed=True).tobytes()
