"""Generates tests/golden/rng_golden.json from the UNMODIFIED reference header
/root/reference/Rtxpt/Shaders/PathTracer/Utils/NoiseAndSequences.hlsli (its C++ half: Hash32, Hash32Combine, Hash32ToFloat, SobolC),
compiled in place by oracle/Makefile into oracle/_ref/ref_kat.  Run in the build container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_rng_golden.py
The committed JSON pins the oracle's and the product's integer sample generators."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_kat")], check=True, capture_output=True, text=True).stdout
data = json.loads(out)
data["_source"] = "Rtxpt/Shaders/PathTracer/Utils/NoiseAndSequences.hlsli (C++ half) at reference commit f08d1c7, via oracle/ref_kat_main.cpp"
with open(os.path.join(ROOT, "tests", "golden", "rng_golden.json"), "w") as f:
    json.dump(data, f, indent=1)
print({k: len(v) for k, v in data.items() if isinstance(v, list)})
