"""Tiny numpy PLY helpers for the tests (the reference ships no data; readme.txt:12)."""
import numpy as np


def write_ply_xyz_rgb(path, xyz_m, rgb, fmt="binary_little_endian", ptype="float"):
    n = len(xyz_m)
    hdr = ["ply", f"format {fmt} 1.0", "comment generated", f"element vertex {n}"]
    hdr += [f"property {ptype} {a}" for a in "xyz"]
    if rgb is not None:
        hdr += ["property uchar red", "property uchar green", "property uchar blue"]
    hdr += ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if fmt == "ascii":
            for i in range(n):
                row = " ".join(repr(float(v)) for v in xyz_m[i])
                if rgb is not None:
                    row += " " + " ".join(str(int(v)) for v in rgb[i])
                f.write((row + "\n").encode())
        else:
            ft = "<f4" if ptype == "float" else "<f8"
            fields = [("x", ft), ("y", ft), ("z", ft)]
            if rgb is not None:
                fields += [("r", "u1"), ("g", "u1"), ("b", "u1")]
            rec = np.zeros(n, dtype=fields)
            rec["x"], rec["y"], rec["z"] = xyz_m[:, 0], xyz_m[:, 1], xyz_m[:, 2]
            if rgb is not None:
                rec["r"], rec["g"], rec["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
            rec.tofile(f)


def read_ply_ref_output(path):
    """Parses what ply::write(binary) produces: float64 xyz + uchar green, blue, red."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode().splitlines()
    n = int([h for h in header if h.startswith("element vertex")][0].split()[2])
    has_col = any("uchar green" in h for h in header)
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if has_col:
        fields += [("g", "u1"), ("b", "u1"), ("r", "u1")]
    rec = np.frombuffer(data[end:], dtype=fields, count=n)
    xyz = np.stack([rec["x"], rec["y"], rec["z"]], 1)
    col = np.stack([rec["g"], rec["b"], rec["r"]], 1) if has_col else None
    return header, xyz, col
