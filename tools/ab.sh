#!/usr/bin/env bash
# tools/ab.sh WORKLOAD NAME...   (on the GPU box): stage table of the default library and of each variants/NAME.so
wl="$1"; shift
run() {
  timeout 200 python bench.py --workload "$wl" --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],3), ' '.join(f\"{k}={v['ms_total']/d['steps']:.3f}\" for k,v in d['stages'].items()))"
}
run default
cp newtonnet_b200/lib/libnewtonnet_b200.so /tmp/default.so
for v in "$@"; do cp "variants/$v.so" newtonnet_b200/lib/libnewtonnet_b200.so; run "$v"; done
cp /tmp/default.so newtonnet_b200/lib/libnewtonnet_b200.so
