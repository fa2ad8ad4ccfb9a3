#!/bin/bash
for v in "base:X=1" "rst4:CTN_TS_RST4=1"; do tag=${v%%:*}; envs=${v#*:}
env $envs timeout 300 python bench.py --config 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$tag c2 bf16', round(d['value']), round(d['ms_per_step'],3))"
done
CTN_TS_RST4=1 timeout 200 python scratch/half_stress.py 2>&1 | tail -1
