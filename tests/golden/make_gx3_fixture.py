"""Generate tests/golden/gx3_grid.npz from the reference's shipped gx3 grid + kmt files.

Run in the build container only (/root/reference is not present on the GPU box):
    python tests/golden/make_gx3_fixture.py
Decodes input_templates/gx3/global_gx3.grid (big-endian direct-access fp64 records:
ULAT, ULON, HTN, HTE, HUS, HUW, ANGLE; /root/reference/source/ice_grid.F90:497-607) and
global_gx3.kmt (big-endian int32).  Known answers: ice.log.Linux.LANL.coyote:101-119.
"""
import os
import numpy as np

REF = "/root/reference/input_templates/gx3"
NX, NY = 100, 116
HERE = os.path.dirname(os.path.abspath(__file__))

raw = np.fromfile(os.path.join(REF, "global_gx3.grid"), dtype=">f8")
assert raw.size == 7 * NX * NY, raw.size
rec = raw.reshape(7, NY, NX).transpose(0, 2, 1)  # -> (rec, i, j)
kmt = np.fromfile(os.path.join(REF, "global_gx3.kmt"), dtype=">i4").reshape(NY, NX).T
names = ["ULAT", "ULON", "HTN", "HTE", "HUS", "HUW", "ANGLE"]
out = {n: np.ascontiguousarray(rec[k]).astype(np.float64) for k, n in enumerate(names)
       if n in ("ULAT", "ULON", "HTN", "HTE")}
out["KMT"] = np.ascontiguousarray(kmt).astype(np.int32)
np.savez_compressed(os.path.join(HERE, "gx3_grid.npz"), **out)
for n in ("ULAT", "HTN", "HTE"):
    print(n, repr(out[n].min()), repr(out[n].max()))
print("KMT max", out["KMT"].max(), "ocean cells", int((out["KMT"] >= 1).sum()))
