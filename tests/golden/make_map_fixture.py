"""Derive ros2_mpc_b200/data/map_carto_occ.npz from the reference's map (run in the build container only).

Reads /root/reference/maps/map_carto.pgm|yaml and applies the pixel convention of
/root/reference/ros2_mpc/core/map_server.py:14-20 (pixel 0 -> occupied, 254/205 -> free, flipud so that row 0
is the bottom of the map).  The GPU box has no /root/reference, so the derived occupancy grid is committed as a
bit-packed fixture together with this script.
"""
import os
import numpy as np
import yaml

REF = "/root/reference/maps"


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    # binary P5: magic, width, height, maxval separated by whitespace (comments allowed)
    toks, i = [], 0
    while len(toks) < 4:
        while data[i:i + 1].isspace():
            i += 1
        if data[i:i + 1] == b"#":
            while data[i:i + 1] != b"\n":
                i += 1
            continue
        j = i
        while not data[j:j + 1].isspace():
            j += 1
        toks.append(data[i:j])
        i = j
    assert toks[0] == b"P5"
    w, h, mx = int(toks[1]), int(toks[2]), int(toks[3])
    assert mx < 256
    img = np.frombuffer(data, dtype=np.uint8, count=w * h, offset=i + 1).reshape(h, w)
    return img


def main():
    img = read_pgm(os.path.join(REF, "map_carto.pgm"))
    with open(os.path.join(REF, "map_carto.yaml")) as f:
        y = yaml.safe_load(f)
    occ = np.flipud(img == 0)  # map_server.py:16,20
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "ros2_mpc_b200", "data", "map_carto_occ.npz")
    free = np.flipud(img == 254)  # known-free pixels (205 = unknown)
    np.savez_compressed(out, occ_bits=np.packbits(occ), free_bits=np.packbits(free), shape=np.array(occ.shape),
                        resolution=float(y["resolution"]), origin=np.array(y["origin"][:2], dtype=np.float64))
    print(out, occ.shape, int(occ.sum()), "occupied cells;", {int(v): int((img == v).sum()) for v in np.unique(img)})


if __name__ == "__main__":
    main()
