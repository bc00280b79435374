# does the e2e figure depend on what ran before it in the process?  (same flags except the stage-timing pass)
for flags in "--no-extras" "--no-extras --no-stage-timing" "--no-extras" "--no-extras --no-stage-timing"; do python bench.py --steps 1000 --no-cpu-baseline $flags 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$flags', '| value', round(d['value']), 'e2e', round(d['e2e']['value']), 'rb32', round(d['e2e_readback_f32']['value']), 'one', round(d['e2e_one_stream']['value']), 'f32', round(d['e2e_f32']['value']))
"; done
