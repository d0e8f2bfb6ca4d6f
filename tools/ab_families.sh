#!/bin/bash
# A/B of per-family builds: tools/ab_families.sh "<workload> <variant.so or ''> ..." pairs are edited in place for an experiment
run() { B200ENV_LIB=$2 python bench.py --workload $1 --steps 100 --warmup 5 --no-extras | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])'; }
V=$PWD/tools/variants
for r in 1 2; do
echo "twolink base $(run twolink '')"; echo "twolink tl10 $(run twolink $V/libb200env_tl10.so)"
echo "ugv base $(run ugv '')"; echo "ugv ugv10 $(run ugv $V/libb200env_ugv10.so)"
done
