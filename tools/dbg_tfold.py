import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from cice4_b200 import evp as E, synth, build
from oracle import oracle as O
from helpers import oracle_steps, cuda_steps
build.build()
case = synth.make_case(name="x", nx=26, ny=20, ew="cyclic", ns="tripoleT", realistic=True)
ndte = 1
st, f, strengths, _ = oracle_steps(O, case, nsteps=1, ndte=ndte)
dyn, out = cuda_steps(case, nsteps=1, strengths=strengths, math_mode=0, ndte=ndte, two_phase=True,
                      want=["strength", "divu", "strintx", "strairx", "prs_sig", "sicemass", "fm"])
np.set_printoptions(linewidth=220, precision=4)
for n in ("icetmask", "strength", "strairx", "fm", "prs_sig", "divu", "strintx"):
    a = out[n][:, :, 0]; b = f[n]
    d = np.argwhere(a != b)
    print(n, "ndiff", len(d), "rows", sorted(set(d[:, 1].tolist())), "cols", sorted(set(d[:, 0].tolist()))[:14])
print("icetmask rows 19..21 cuda/oracle")
for j in (19, 20, 21):
    print(j, out["icetmask"][:, j, 0]); print(j, f["icetmask"][:, j])
print("strength rows 20,21 cuda/oracle")
for j in (20, 21):
    print(j, out["strength"][:, j, 0]); print(j, f["strength"][:, j])
print("tmask / hm top rows", case.grid.f["tmask"][:, 19:22].T)
