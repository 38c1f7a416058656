"""Harness-side reader / writer for DL_POLY CONFIG files.

Format as written by the reference (centroid_write_config.f:38-87; same as DL_POLY 2 CONFIG / REVCON):
    line 1   title                                   (a80)
    line 2   levcfg, imcon, nstep, t                 (3i10, es25.16)   -- classic CONFIGs carry only levcfg, imcon
    3 lines  cell(1:9), three numbers per line        (3es25.16)        -- only when imcon > 0
    per atom: name, index                            (a8, i10)
              x y z                                  (3es25.16)
              vx vy vz     if levcfg >= 1
              fx fy fz     if levcfg >= 2
Lets equilibrated MB-pol boxes be fed to bench.py / the tests instead of the synthetic lattice.  Waters are
recognised as consecutive (O*, H*, H*) triples, exactly what ttm2list.f hands to mbpol_forces as listttm2.
"""
import numpy as np


def read_config(path):
    with open(path) as f:
        lines = f.read().splitlines()
    title = lines[0]
    head = lines[1].split()
    levcfg, imcon = int(head[0]), int(head[1])
    pos = 2
    cell = np.zeros(9)
    if imcon > 0:
        cell = np.array([float(v) for k in range(3) for v in lines[pos + k].split()[:3]])
        pos += 3
    names, xyz, vel, frc = [], [], [], []
    while pos < len(lines) and lines[pos].strip():
        names.append(lines[pos][:8].strip())
        xyz.append([float(v) for v in lines[pos + 1].split()[:3]])
        pos += 2
        if levcfg >= 1:
            vel.append([float(v) for v in lines[pos].split()[:3]])
            pos += 1
        if levcfg >= 2:
            frc.append([float(v) for v in lines[pos].split()[:3]])
            pos += 1
    xyz = np.array(xyz)
    n = len(names)
    vel = np.array(vel) if vel else np.zeros((n, 3))
    lst = []
    i = 0
    while i + 2 < n:
        if names[i][:1].upper() == "O" and names[i + 1][:1].upper() == "H" and names[i + 2][:1].upper() == "H":
            lst += [i + 1, i + 2, i + 3]
            i += 3
        else:
            i += 1
    box = {"title": title, "levcfg": levcfg, "imcon": imcon, "cell": cell, "names": names, "mxatms": n, "nw": len(lst) // 3,
           "nttm2": len(lst), "listttm2": np.array(lst, dtype=np.int32), "L": float(cell[0]), "seed": -1,
           "xxx": np.ascontiguousarray(xyz[:, 0]), "yyy": np.ascontiguousarray(xyz[:, 1]), "zzz": np.ascontiguousarray(xyz[:, 2]),
           "vxx": np.ascontiguousarray(vel[:, 0]), "vyy": np.ascontiguousarray(vel[:, 1]), "vzz": np.ascontiguousarray(vel[:, 2])}
    if frc:
        f = np.array(frc)
        box["fxx"], box["fyy"], box["fzz"] = (np.ascontiguousarray(f[:, k]) for k in range(3))
    return box


def write_config(path, box, forces=None, title="mbpol_b200 water box", nstep=0, t=0.0):
    n = int(box["mxatms"])
    nw = int(box["nw"])
    names = box.get("names") or (["OW", "HW", "HW"] * nw + ["MW"] * (n - 3 * nw))
    zero = np.zeros(n)
    f = forces if forces is not None else (zero, zero, zero)

    def e3(a, b, c):
        return "%25.16E%25.16E%25.16E" % (a, b, c)

    with open(path, "w") as out:
        out.write("%-80s\n" % title[:80])
        out.write("%10d%10d%10d%25.16E\n" % (2, int(box["imcon"]), nstep, t))
        if int(box["imcon"]) > 0:
            c = box["cell"]
            for k in range(3):
                out.write(e3(c[3 * k], c[3 * k + 1], c[3 * k + 2]) + "\n")
        for i in range(n):
            out.write("%-8s%10d\n" % (names[i], i + 1))
            out.write(e3(box["xxx"][i], box["yyy"][i], box["zzz"][i]) + "\n")
            out.write(e3(box["vxx"][i], box["vyy"][i], box["vzz"][i]) + "\n")
            out.write(e3(f[0][i], f[1][i], f[2][i]) + "\n")


def read_field(path):
    """Minimal DL_POLY 2 FIELD reader: what the MB-pol path needs from it is the molecular make-up -- molecule types with `nummols` copies
    of `numsit` sites each (sysdef.f) -- from which ttm2list.f:27-70 derives listttm2: the water is the SECOND TO LAST molecule type (3 sites),
    the last type holds the M sites (1 site, as many as waters), atoms are numbered type by type, molecule by molecule, site by site, and the
    heaviest of the three water sites (the oxygen) is moved to the first place of every triple.
    Returns {"types": [(name, nummols, [(site name, mass, charge), ...])], "natoms", "nttm2", "listttm2" (1-based), "names"}."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    types, i, ntp = [], 0, None
    while i < len(lines):
        key = lines[i].strip().lower()
        if key.startswith("molecul") and ntp is None:          # "MOLECULES n" / "molecular types n"
            ntp = int(key.split()[-1])
            i += 1
            for _ in range(ntp):
                name = lines[i].strip()
                i += 1
                nummols, sites = 1, []
                while not lines[i].strip().lower().startswith("finish"):
                    w = lines[i].split()
                    k = w[0].lower()
                    if k.startswith("nummol"):
                        nummols = int(w[1])
                        i += 1
                    elif k.startswith("atoms"):
                        na = int(w[1])
                        i += 1
                        while len(sites) < na:
                            a = lines[i].split()
                            rep = int(a[3]) if len(a) > 3 and a[3].lstrip("-").isdigit() else 1
                            sites += [(a[0], float(a[1]), float(a[2]))] * max(rep, 1)
                            i += 1
                    else:                                     # bonds, angles, constraints, rigid units ...: "<keyword> n" + n lines
                        n = int(w[1]) if len(w) > 1 and w[1].isdigit() else 0
                        i += 1 + n
                i += 1
                types.append((name, nummols, sites))
        else:
            i += 1
    if ntp is None:
        raise ValueError("no MOLECULES directive in %s" % path)
    names = [s[0] for (_, nm, sites) in types for _ in range(nm) for s in sites]
    out = {"types": types, "natoms": len(names), "names": names, "nttm2": 0, "listttm2": np.zeros(0, dtype=np.int32)}
    if len(types) >= 2 and len(types[-2][2]) == 3 and len(types[-1][2]) == 1 and types[-2][1] == types[-1][1]:      # ttm2list.f:19-30
        first = sum(nm * len(s) for (_, nm, s) in types[:-2])
        nw = types[-2][1]
        lst = np.arange(first + 1, first + 3 * nw + 1, dtype=np.int32).reshape(nw, 3)
        ox = int(np.argmax([s[1] for s in types[-2][2]]))                                                       # ttm2list.f:54-61
        if ox != 0:
            lst[:, [0, ox]] = lst[:, [ox, 0]]
        out["nttm2"], out["listttm2"] = 3 * nw, lst.ravel().copy()
    return out


def write_field(path, box, title="mbpol_b200 water box"):
    """FIELD of a synth box: [ions] + `nw` TTM/MB-pol waters (O, H, H) + `nw` M sites, the layout ttm2list.f expects."""
    nw = int(box["nw"])
    nions = int(box.get("nions", 0))
    with open(path, "w") as out:
        out.write("%s\nUNITS kcal\nMOLECULES %d\n" % (title, 2 + (1 if nions else 0)))
        if nions:
            out.write("ions\nNUMMOLS %d\nATOMS 1\nNa+ 22.9898 1.0\nFINISH\n" % nions)
        out.write("water\nNUMMOLS %d\nATOMS 3\nOW 15.9994 0.0\nHW 1.008 0.0\nHW 1.008 0.0\nFINISH\n" % nw)
        out.write("msite\nNUMMOLS %d\nATOMS 1\nMW 0.0 0.0\nFINISH\nVDW 0\nCLOSE\n" % nw)
