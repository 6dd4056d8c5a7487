#!/bin/bash
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, os, json
sys.path.insert(0, 'scripts')
import bench_conv
for s in [(128, 32, 16, 16), (128, 16, 32, 32), (128, 8, 64, 64)]:
    print('nstack', s, bench_conv.bench(*s))
PY
SE_WG_NO_NSTACK=1 python - <<'PY'
import sys, os, json
sys.path.insert(0, 'scripts')
import bench_conv
for s in [(128, 32, 16, 16), (128, 16, 32, 32), (128, 8, 64, 64)]:
    print('no-nstack', s, bench_conv.bench(*s))
PY
SE_WG_DEBUG=4 python - <<'PY'
import sys, os, json
sys.path.insert(0, 'scripts')
import bench_conv
for s in [(128, 32, 16, 16), (128, 16, 32, 32)]:
    print('no-mma', s, bench_conv.bench(*s))
PY
