#!/bin/bash
# A/B of compile-time variants of the library on the GPU box (measurement tool, not a test):
#   built HERE (no GPU needed):   tests/ab_variants.sh build NAME -DRS_VPW=1 ...      -> tests/_build/variants/NAME.so
#   run on the box (gpurun):      tests/ab_variants.sh run NAME [NAME ...]            -> gpurun_out/variants.jsonl
# Each variant is timed by tests/ab_toggles.py's baseline leg (device-resident frame, host-pointer frame, per-kernel times).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
V="$ROOT/tests/_build/variants"
case "$1" in
build) shift; name="$1"; shift; mkdir -p "$V"; python "$ROOT/dsp-map_b200/build.py" --out "$V/$name.so" "$@" ;;
run) shift
  for name in "$@"; do
    lib="$V/$name.so"; [ "$name" = default ] && lib="$ROOT/dsp-map_b200/lib/libdspmap_b200.so"
    DSPMAP_B200_LIB="$lib" timeout 300 python "$ROOT/tests/ab_toggles.py" --inproc --frames 0 --steps 30 --baseline measure --out "$ROOT/gpurun_out/variant_$name.jsonl" BASELINE > "$ROOT/gpurun_out/variant_$name.log" 2>&1 || echo "$name failed"
    python - "$name" "$ROOT/gpurun_out/variant_$name.jsonl.baseline" <<'PY'
import json, sys
b = json.load(open(sys.argv[2]))
top = sorted(b["kernels"].items(), key=lambda kv: -kv[1])[:12]
print(json.dumps({"variant": sys.argv[1], "device_ms": round(b["dev"], 4), "host_api_ms": round(b["host"], 4), "kernels_us": dict(top)}))
PY
  done ;;
esac
