import numpy as np


def test_config_round_trip(tmp_path, synth):
    """CONFIG writer/reader (centroid_write_config.f:38-87 format): lossless at es25.16, waters recognised as O,H,H triples."""
    from mbpol_b200 import config_io
    box = synth.make_water_box(27, seed=5)
    p = tmp_path / "CONFIG"
    config_io.write_config(str(p), box)
    txt = p.read_text().splitlines()
    assert len(txt) == 2 + 3 + 4 * box["mxatms"] and txt[1].split()[:2] == ["2", "1"]
    b2 = config_io.read_config(str(p))
    assert b2["nw"] == 27 and b2["mxatms"] == box["mxatms"] and b2["imcon"] == 1
    assert np.array_equal(b2["listttm2"], box["listttm2"])
    for k in ("xxx", "yyy", "zzz", "vxx", "vyy", "vzz", "cell"):
        np.testing.assert_allclose(b2[k], box[k], rtol=1e-15, atol=1e-300)


def test_field_reader_gives_ttm2list(tmp_path, synth):
    """read_field reproduces ttm2list.f: the second-to-last molecule type is the water, numbered after everything before it, oxygen first."""
    from mbpol_b200 import config_io
    box = synth.make_water_box(64)
    p = tmp_path / "FIELD"
    config_io.write_field(str(p), box)
    f = config_io.read_field(str(p))
    assert f["natoms"] == box["mxatms"] and f["nttm2"] == box["nttm2"]
    assert np.array_equal(f["listttm2"], box["listttm2"])
    # hydrogen listed first in the FIELD: the oxygen (heaviest site) is still moved to the front of every triple
    p.write_text("t\nUNITS kcal\nMOLECULES 3\nions\nNUMMOLS 2\nATOMS 1\nCl- 35.45 -1.0\nFINISH\nwater\nNUMMOLS 3\nATOMS 3\nHW 1.008 0.0\n"
                 "OW 15.9994 0.0\nHW 1.008 0.0\nBONDS 2\nharm 1 2 1.0 1.0\nharm 2 3 1.0 1.0\nFINISH\nmsite\nNUMMOLS 3\nATOMS 1\nMW 0.0 0.0\nFINISH\nCLOSE\n")
    f = config_io.read_field(str(p))
    assert f["natoms"] == 2 + 9 + 3 and f["listttm2"].tolist() == [4, 3, 5, 7, 6, 8, 10, 9, 11]
