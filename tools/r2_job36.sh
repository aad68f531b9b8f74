cd $GRAFT_REPO_ROOT
for c in 96 128 148 0; do
echo "TC_CTAS=$c"; B200IPM_LDLT_TC_CTAS=$c timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('native factor', round(d['native']['factor_ms'],2), 'solve', round(d['native']['solve_ms_8rhs_1refine'],2), 'pipeline', round(d['pipeline_on_one_rank']['factor_ms'],2), d['native']['scaled_residual_inf'])"
done
