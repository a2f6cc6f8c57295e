#!/bin/bash
# run bench.py for the default library and every variant under merzbild.jl_b200/_variants; one summary line each
# usage: profiles/run_variants.sh "<bench args>" [variant names...]
args=$1; shift
names=${@:-default $(ls merzbild.jl_b200/_variants 2>/dev/null | sed 's/libmb_//; s/.so//')}
for v in $names; do
  if [ "$v" = default ]; then unset MERZBILD_B200_LIB; else export MERZBILD_B200_LIB=$PWD/merzbild.jl_b200/_variants/libmb_$v.so; fi
  python bench.py $args --no-cpu-baseline --no-others --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', '%.3f ms' % d['ms_per_step'], 'frac %.3f' % (d['roofline']['frac'] or 0), {k: round(x, 3) for k, x in d['roofline']['sections_ms_per_step'].items()}, d['config'].get('sort_path'))"
done
