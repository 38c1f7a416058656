"""Synthetic MB-pol water boxes (the reference ships no CONFIG/FIELD inputs).

Layout mirrors what mbpol_forces receives from DL_POLY (dlpoly_cmd.f:2504, ttm2list.f): SoA coordinate
arrays xxx/yyy/zzz of length mxatms = 4*nw with the atoms of water m at 1-based indices 3m-2 (O), 3m-1, 3m (H)
followed by nw M-sites, listttm2 = 1..3nw, cubic cell (imcon = 1), velocities from a 298 K Maxwell-Boltzmann.
"""
import numpy as np

RHO = 0.03343           # molecules / A^3
R_OH = 0.9572           # A
A_HOH = 104.52          # degrees
SEEDS = {256: 1001, 4096: 1002, 32768: 1003, 512: 1005}


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def make_water_box(nw, seed=None, min_contact=1.5, max_sweeps=20):
    """Jittered simple-cubic lattice of rigid, randomly oriented waters; orientations are re-drawn until no
    intermolecular atom-atom distance is below `min_contact` (a handful of close contacts would otherwise
    dominate the energy)."""
    from scipy.spatial import cKDTree

    if seed is None:
        seed = SEEDS.get(nw, 1000 + nw)
    rng = np.random.default_rng(seed)
    L = (nw / RHO) ** (1.0 / 3.0)
    m = int(round(nw ** (1.0 / 3.0)))
    while m ** 3 < nw:
        m += 1
    a = L / m
    grid = np.stack(np.meshgrid(*(np.arange(m),) * 3, indexing="ij"), -1).reshape(-1, 3)[:nw]
    O = (grid + 0.5) * a - L / 2 + rng.uniform(-0.25, 0.25, size=(nw, 3))
    th = np.deg2rad(A_HOH) / 2
    h_local = np.array([[R_OH * np.sin(th), 0.0, R_OH * np.cos(th)], [-R_OH * np.sin(th), 0.0, R_OH * np.cos(th)]])

    def place(idx):
        R = _random_rotations(rng, len(idx))
        return O[idx, None, :] + np.einsum("nij,kj->nki", R, h_local)

    H = place(np.arange(nw))
    for _ in range(max_sweeps):
        atoms = np.concatenate([O[:, None, :], H], axis=1).reshape(-1, 3)
        wrapped = (atoms + L / 2) % L
        tree = cKDTree(wrapped, boxsize=L)
        pairs = tree.query_pairs(min_contact, output_type="ndarray")
        if len(pairs):
            pairs = pairs[pairs[:, 0] // 3 != pairs[:, 1] // 3]
        if not len(pairs):
            break
        bad = np.unique(pairs // 3)
        H[bad] = place(bad)
    xyz = np.concatenate([O[:, None, :], H], axis=1) + rng.normal(scale=0.02, size=(nw, 3, 3))
    xyz = xyz - L * np.round(xyz / L)          # wrap every atom into the primary cell (molecules may be split)

    mx = 4 * nw
    pos = np.zeros((mx, 3))
    pos[:3 * nw] = xyz.reshape(-1, 3)
    # M-sites (not used by MB-pol 2B/3B; present so that mxatms = 4 nw like a TTM/MB-pol FIELD)
    oh1 = xyz[:, 1] - xyz[:, 0]; oh1 -= L * np.round(oh1 / L)
    oh2 = xyz[:, 2] - xyz[:, 0]; oh2 -= L * np.round(oh2 / L)
    pos[3 * nw:] = xyz[:, 0] + 0.2 * 0.5 * (oh1 + oh2)
    kT = 0.831451115 * 298.0                    # DL_POLY internal units (10 J/mol): boltz = 0.831451115
    mass = np.tile([15.9994, 1.008, 1.008], nw)
    vel = np.zeros((mx, 3))
    vel[:3 * nw] = rng.normal(size=(3 * nw, 3)) * np.sqrt(kT / mass)[:, None]
    cell = np.zeros(9)
    cell[0] = cell[4] = cell[8] = L
    return {
        "nw": nw, "mxatms": mx, "nttm2": 3 * nw, "imcon": 1, "cell": cell, "L": L, "seed": seed,
        "listttm2": np.arange(1, 3 * nw + 1, dtype=np.int32),
        "xxx": np.ascontiguousarray(pos[:, 0]), "yyy": np.ascontiguousarray(pos[:, 1]), "zzz": np.ascontiguousarray(pos[:, 2]),
        "vxx": np.ascontiguousarray(vel[:, 0]), "vyy": np.ascontiguousarray(vel[:, 1]), "vzz": np.ascontiguousarray(vel[:, 2]),
    }


def make_ion_box(nw_total=4096, nions=64, ions=("na", "cl"), seed=1004, min_contact=2.0):
    """BASELINE config 4: the water box of make_water_box(nw_total) in which `nions` randomly chosen
    waters are replaced by ions (alternating species, placed at the removed oxygen).  Ions are the FIRST atoms, as
    mbnrg.f90:226-233 requires; then O,H,H per water; then the M-sites.  Waters closer than `min_contact` to an ion are
    not removed (the synthetic lattice spacing of 3.1 A keeps ion-oxygen distances physical)."""
    from .mbnrg import ION_CODES
    base = make_water_box(nw_total, seed=seed)
    rng = np.random.default_rng(seed + 17)
    gone = np.sort(rng.choice(nw_total, size=nions, replace=False))
    keep = np.setdiff1d(np.arange(nw_total), gone)
    nw = len(keep)
    pos_w = np.stack([base["xxx"], base["yyy"], base["zzz"]], 1)[:3 * nw_total].reshape(nw_total, 3, 3)
    vel_w = np.stack([base["vxx"], base["vyy"], base["vzz"]], 1)[:3 * nw_total].reshape(nw_total, 3, 3)
    mx = nions + 4 * nw
    pos, vel = np.zeros((mx, 3)), np.zeros((mx, 3))
    pos[:nions] = pos_w[gone, 0]
    pos[nions:nions + 3 * nw] = pos_w[keep].reshape(-1, 3)
    vel[nions:nions + 3 * nw] = vel_w[keep].reshape(-1, 3)
    code = np.array([ION_CODES[ions[j % len(ions)]] for j in range(nions)], dtype=np.int32)
    return {
        "nw": nw, "nions": nions, "mxatms": mx, "nttm2": 3 * nw, "imcon": 1, "cell": base["cell"], "L": base["L"], "seed": seed,
        "listttm2": np.arange(nions + 1, nions + 3 * nw + 1, dtype=np.int32), "listions": np.arange(1, nions + 1, dtype=np.int32),
        "ion_code": code,
        "xxx": np.ascontiguousarray(pos[:, 0]), "yyy": np.ascontiguousarray(pos[:, 1]), "zzz": np.ascontiguousarray(pos[:, 2]),
        "vxx": np.ascontiguousarray(vel[:, 0]), "vyy": np.ascontiguousarray(vel[:, 1]), "vzz": np.ascontiguousarray(vel[:, 2]),
    }
