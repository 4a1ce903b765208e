#!/bin/bash
# same-box A/B of environment toggles: scripts/exp/ab_env.sh "VAR=val VAR2=val" ... (each argument is one variant; "" = defaults)
run() { env $1 python bench.py --no-extras --skip-cpu-baseline --skip-eager --skip-profile --steps 40 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-44s %.4f ms  e2e %.4f ms' % ('${1:-defaults}', d['ms_per_step'], d['e2e']['ms_per_step']))"; }
for v in "$@"; do run "$v"; done
