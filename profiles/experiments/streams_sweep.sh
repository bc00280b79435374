# feed_data calls in flight (--streams) against distinct input sets (OTF_BENCH_ROTATE); prints value, bare graph replay, e2e
for cfg in "4 4 0" "8 8 0" "8 8 1" "4 4 1"; do set -- $cfg; if [ "$3" = "1" ]; then export OTF_PARAM_PINNED_EXPERIMENT=1; else unset OTF_PARAM_PINNED_EXPERIMENT; fi; OTF_BENCH_ROTATE=$1 python bench.py --steps 1000 --streams $2 --no-cpu-baseline --no-extras --no-stage-timing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rotate/streams/pinned', '$cfg', round(d['value']), round(d['value_graph_replay']), round(d['e2e']['value']))
"; done
