"""Parse the real geometries the reference ships (raw/AID_kcal.xyz: 451 molecules of 60-146 atoms, the only
usable coordinates in the mount; format handled by utils.py:17-63 `read_xyz`: atom count, label line, then
`element x y z` rows) into a small numeric fixture for the "OCELOT-sized" configuration (BASELINE.json
configs[2]).  Only numbers are stored: atomic numbers, positions, per-molecule
offsets and the energy label.

    python tests/golden/make_aid_fixture.py        (build container only; writes tests/golden/aid_geometries.npz)
"""
import os

import numpy as np

SRC = "/root/reference/raw/AID_kcal.xyz"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "aid_geometries.npz")
Z_OF = {"H": 1, "C": 6, "N": 7, "O": 8, "F": 9}        # utils.py:19


def main():
    zs, pos, ptr, labels = [], [], [0], []
    with open(SRC) as f:
        lines = [ln.split() for ln in f if ln.strip()]
    i = 0
    while i < len(lines):
        n = int(lines[i][0])
        labels.append(float(lines[i + 1][0]))
        for el, x, y, z in lines[i + 2:i + 2 + n]:
            zs.append(Z_OF[el])
            pos.append((float(x), float(y), float(z)))
        ptr.append(len(zs))
        i += 2 + n
    np.savez_compressed(DST, z=np.asarray(zs, np.uint8), pos=np.asarray(pos, np.float32),
                        ptr=np.asarray(ptr, np.int32), label=np.asarray(labels, np.float64))
    sizes = np.diff(ptr)
    print(f"{len(sizes)} molecules, {len(zs)} atoms, sizes {sizes.min()}..{sizes.max()} (mean {sizes.mean():.1f}), "
          f"{os.path.getsize(DST) // 1024} KiB -> {DST}")


if __name__ == "__main__":
    main()
