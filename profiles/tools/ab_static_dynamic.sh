for i in 1 2 3; do
for v in dyn static; do
if [ $v = static ]; then export DODO_PRUNE_STATIC=1; else unset DODO_PRUNE_STATIC; fi
python bench.py --steps 20 --warmup 5 --no-strong --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['clocks']['sm_mhz'])"
done; done
