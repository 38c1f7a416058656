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
