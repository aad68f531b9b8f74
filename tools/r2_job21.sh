cd $GRAFT_REPO_ROOT
for c in 32 48 64 96 128 148; do
  echo "TC_CTAS=$c"; B200IPM_LDLT_TC_CTAS=$c timeout 300 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|periods"
done
echo "TC=0"; B200IPM_LDLT_TC=0 timeout 300 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|periods"
