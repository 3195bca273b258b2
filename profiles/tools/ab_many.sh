#!/bin/bash
# ab_many.sh reps lib1.so lib2.so ...: the headline step with several builds of the library, round-robin, on ONE box
reps=$1; shift
for i in $(seq $reps); do
for v in "$@"; do
DODO_LIB=$v python bench.py --steps 20 --warmup 5 --no-strong --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['clocks']['sm_mhz'])"
done; done
