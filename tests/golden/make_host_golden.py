"""Generates tests/golden/host_golden.json from the UNMODIFIED reference headers / sources compiled in place by oracle/Makefile into
oracle/_ref/ref_kat_host (PathTracerShared.h BridgeCamera + structs, MaterialPT.h, SubInstanceData.h, PolymorphicLight.h, Donut bindless.h and
core/math/vector.cpp).  Run in the build container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_host_golden.py
The committed JSON pins the C ABI struct mirrors, the material / light constants, the camera bridge and the snorm8 vertex packing."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_kat_host")], check=True, capture_output=True, text=True).stdout
data = json.loads(out)
data["_source"] = ("Rtxpt/Shaders/PathTracer/PathTracerShared.h, Materials/MaterialPT.h, SubInstanceData.h, Lighting/PolymorphicLight.h, External/Donut/include/donut/shaders/bindless.h, "
                   "External/Donut/src/core/math/vector.cpp at reference commit f08d1c7, via oracle/ref_kat_host_main.cpp")
with open(os.path.join(ROOT, "tests", "golden", "host_golden.json"), "w") as f:
    json.dump(data, f, indent=1)
print({k: (v["size"] if isinstance(v, dict) and "size" in v else len(v)) for k, v in data.items() if not k.startswith("_")})
