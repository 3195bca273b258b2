#!/bin/bash
# ab_variants.sh A.so B.so [reps]: the headline step with two builds of the library, alternating, on ONE GPU box
# (box-to-box differences are ~1 %: only numbers from the same box compare)
reps=${3:-3}
for i in $(seq $reps); do
for v in "$1" "$2"; do
DODO_LIB=$v python bench.py --steps 20 --warmup 5 --no-strong --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['clocks']['sm_mhz'])"
done; done
