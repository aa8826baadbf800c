"""Writes the inputs of the small golden case (tests/golden/decon_small.npz, conv_small.npz) as raw little-endian float32
files (x fastest, then y, then z -- ImgLib's and the JNA boundary's order) for GenerateGolden.java:

    python tests/golden/reference/export_inputs.py          ->  tests/golden/reference/inputs/

    meta.txt            nx ny nz num_views kx ky kz  /  conv: nx ny nz kx ky kz
    img<v>.raw  w<v>.raw  psf<v>.raw  conv_img.raw  conv_kernel.raw
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.dirname(HERE)


def main():
    out = os.path.join(HERE, "inputs")
    os.makedirs(out, exist_ok=True)
    d = np.load(os.path.join(G, "decon_small.npz"))
    V = int(d["num_views"])
    nz, ny, nx = (int(x) for x in d["shape"])
    kz, ky, kx = d["psf0"].shape
    for v in range(V):
        d[f"img{v}"].astype("<f4").tofile(os.path.join(out, f"img{v}.raw"))
        d[f"w{v}"].astype("<f4").tofile(os.path.join(out, f"w{v}.raw"))
        d[f"psf{v}"].astype("<f4").tofile(os.path.join(out, f"psf{v}.raw"))
    c = np.load(os.path.join(G, "conv_small.npz"))
    c["img"].astype("<f4").tofile(os.path.join(out, "conv_img.raw"))
    c["kernel"].astype("<f4").tofile(os.path.join(out, "conv_kernel.raw"))
    cz, cy, cx = c["img"].shape
    qz, qy, qx = c["kernel"].shape
    with open(os.path.join(out, "meta.txt"), "w") as f:
        f.write(f"{nx} {ny} {nz} {V} {kx} {ky} {kz}\n{cx} {cy} {cz} {qx} {qy} {qz}\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
