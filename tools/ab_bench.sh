#!/bin/bash
# runs bench.py (device-resident, no extras) for the default library and every tools/variants/*.so
wl=${1:-uav_pos}; steps=${2:-100}
echo "base $(python bench.py --workload $wl --steps $steps --warmup 5 --no-extras | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])')"
for f in tools/variants/*.so; do
  echo "$(basename $f) $(B200ENV_LIB=$PWD/$f python bench.py --workload $wl --steps $steps --warmup 5 --no-extras | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])')"
done
