for cfg in "128 4" "64 7" "64 8" "96 5" "32 7"; do set -- $cfg; MB_MERGE_THREADS=$1 MB_MERGE_PER_SM=$2 python bench.py --config c2 --steps 12 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg', '%.2f ms/step' % d['ms_per_step'], 'merge %.2f' % d['roofline']['sections_ms_per_step']['merge'], 'initial %.1f ms' % d['config']['initial_merge']['ms'], d['config']['merges_in_timed_steps']['cells'])"; done
