import numpy as np


def test_box_layout(synth):
    b = synth.make_water_box(64)
    assert b["mxatms"] == 256 and b["nttm2"] == 192 and b["imcon"] == 1
    assert np.array_equal(b["listttm2"], np.arange(1, 193))
    L = b["L"]
    assert abs(L - (64 / 0.03343) ** (1 / 3)) < 1e-12
    for k in ("xxx", "yyy", "zzz"):
        assert np.all(np.abs(b[k][:192]) <= L / 2 + 1e-9)
    b2 = synth.make_water_box(64)
    assert np.array_equal(b["xxx"], b2["xxx"])            # deterministic
    # rigid-ish water geometry after imaging H next to O
    pos = np.stack([b["xxx"], b["yyy"], b["zzz"]], 1)[:192].reshape(64, 3, 3)
    d = pos[:, 1] - pos[:, 0]
    d -= L * np.round(d / L)
    r = np.linalg.norm(d, axis=1)
    assert np.all(np.abs(r - 0.9572) < 0.15)
