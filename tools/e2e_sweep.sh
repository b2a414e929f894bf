# end-to-end leg (pinned host buffers through ShardedHostEnv) for several slice layouts
run() { echo -n "$1: "; python bench.py --steps 100 --warmup 10 --no-cpu --no-vecenv --transition-envs 0 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print(round(e['value']/1e6,3), 'single', round(e['single_handle_value']/1e6,3), 'bound', round(e['copy_bound_value']/1e6,2), 'value', round(d['value']/1e6,2))"; }
run "4 even" "--host-shards 4"
run "8 even" "--host-shards 8"
run "30,30,25,12,3" "--host-shard-weights 30,30,25,12,3"
run "28,26,22,14,7,3" "--host-shard-weights 28,26,22,14,7,3"
run "40,30,18,8,3,1" "--host-shard-weights 40,30,18,8,3,1"
run "50,30,15,5" "--host-shard-weights 50,30,15,5"
