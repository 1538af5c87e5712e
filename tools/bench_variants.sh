#!/bin/bash
# bench.py (config 2, no sub-configs, no CPU leg) for the default library and every tools/variants/* build
for v in "" tools/variants/*/libmouse_b200.so; do
  echo "== ${v:-default}"
  MOUSE_B200_LIB=$v MOUSE_BENCH_DIAG=1 python bench.py --steps ${STEPS:-200} --warmup 5 --no-configs --no-cpu-baseline 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.3e ms/step %.4f e2e %.4f serial %.4f sweep %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["serial_frame_ms"], d["roofline"]["sweep_ms"], d["roofline"]["frac"]))
    elif "overlapped" in l: print(l.strip()[:200])
'
done
