#!/bin/bash
run() { lab=$1; shift
  for c in 1 3; do env "$@" timeout 200 python bench.py --config $c --no-cpu-baseline --steps 10 2>/dev/null | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$lab c$c', round(d['value']), round(d['ms_per_step'],3), round(d['fwd']['value']))"; done
}
run default X=1
run ncap256 CTN_TS_NCAP=256
run ncap208 CTN_TS_NCAP=208
