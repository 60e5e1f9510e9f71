#!/usr/bin/env python3
"""Extract the two *data tables* the render path needs from the reference into binary blobs.

These are published data sets, not algorithms:
  * the 128x128 RG8 blue-noise tile  (reference: src/pt/blue_noise.c:3, dims :1730-1731)
  * the Hosek-Wilkie RGB sky-model coefficient / radiance tables and the reference author's
    integrated solar radiances (reference: src/hw-skymodel/params_{r,g,b}.h,
    radiances_{r,g,b}.h), consumed by sky_state_new (hw_skymodel.c:141-180).

Outputs (committed, little-endian):
  rayfinder_b200/data/blue_noise_128x128_rg8.bin   32768 x u8
  rayfinder_b200/data/hw_sky_rgb_tables.bin        f32: for c in r,g,b: params[1080], radiances[120],
                                                   solar_radiances[10]   (3630 floats)
Run once in the build container (needs /root/reference); the GPU box only sees the blobs.
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "rayfinder_b200" / "data"


def c_array(text: str, name: str) -> list[str]:
    m = re.search(name + r"\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}", text, re.S)
    assert m, name
    return [t for t in re.split(r"[\s,]+", m.group(1)) if t]


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    bn = c_array((REF / "src/pt/blue_noise.c").read_text(), "blueNoiseValues")
    bn = np.array([int(t) for t in bn], dtype=np.uint8)
    assert bn.size == 128 * 128 * 2
    bn.tofile(OUT / "blue_noise_128x128_rg8.bin")

    blobs = []
    for ch in "rgb":
        ptxt = (REF / f"src/hw-skymodel/params_{ch}.h").read_text()
        rtxt = (REF / f"src/hw-skymodel/radiances_{ch}.h").read_text()
        params = np.array([np.float32(t.rstrip("f")) for t in c_array(ptxt, f"params_{ch}")], dtype=np.float32)
        rad = np.array([np.float32(t.rstrip("f")) for t in c_array(rtxt, f"const float radiances_{ch}")], dtype=np.float32)
        sol = np.array([np.float32(t.rstrip("f")) for t in c_array(rtxt, f"solar_radiances_{ch}")], dtype=np.float32)
        assert params.size == 2 * 10 * 6 * 9 and rad.size == 2 * 10 * 6 and sol.size == 10, (params.size, rad.size, sol.size)
        blobs += [params, rad, sol]
    tables = np.concatenate(blobs).astype("<f4")
    assert tables.size == 3 * (1080 + 120 + 10)
    tables.tofile(OUT / "hw_sky_rgb_tables.bin")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
