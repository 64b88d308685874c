"""Build variants of the library with different -D tuning macros (experiments only):
    python tools/build_variants.py NAME=-DFOO=1,-DBAR=2 ...   ->  sgp_b200/variants/libsgp_b200_NAME.so
Select one at run time with SGP_B200_SO=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgp_b200 import _build
out_dir = os.path.join(_build.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    flags = [f for f in flags.split(",") if f]
    objs = []
    procs = []
    for src in _build.SOURCES:
        obj = os.path.join(out_dir, f"{name}_{src.replace('.cu', '.o')}")
        procs.append(subprocess.Popen([_build._nvcc(), *_build.NVCC_FLAGS, *flags, "-c", os.path.join(_build.CSRC, src), "-o", obj]))
        objs.append(obj)
    assert all(p.wait() == 0 for p in procs)
    so = os.path.join(out_dir, f"libsgp_b200_{name}.so")
    subprocess.check_call([_build._nvcc(), "-shared", "-o", so, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    for o in objs:
        os.remove(o)
    print(so)
