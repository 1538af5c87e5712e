#!/bin/bash
# bench.py (config 2 only) over library variants x frames in flight: tools/bench_warps.sh
for v in "" tools/variants/*/libmouse_b200.so; do
for l in 3 4; do
MOUSE_B200_LIB=$v MOUSE_BENCH_LANES=$l python bench.py --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('${v:-default} lanes $l: value %.4e ms %.4f e2e %.4f serial %.4f sweep %.4f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['serial_frame_ms'], d['roofline']['sweep_ms']))
"
done; done
