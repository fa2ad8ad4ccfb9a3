#!/bin/bash
run() { # label, env...
  lab=$1; shift
  for c in 4 2; do env "$@" timeout 200 python bench.py --config $c --no-cpu-baseline --steps 10 2>/dev/null | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$lab c$c', round(d['value']), round(d['ms_per_step'],3))"; done
}
run default X=1
run ncap128 CTN_TS_NCAP=128
run ncap192 CTN_TS_NCAP=192
run ncap256 CTN_TS_NCAP=256
run rst4 CTN_TS_RST4=1
run ncap256_rst4 CTN_TS_NCAP=256 CTN_TS_RST4=1
run cl1 CTN_TS_CL=1
run cl4 CTN_TS_CL=4
