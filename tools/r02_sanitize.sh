#!/bin/bash
# compute-sanitizer over one small invocation of every kernel family (tools/sanitize.py).  Outputs: gpurun_out/r02s_*
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
mkdir -p $O
python tools/sanitize.py > $O/r02s_sanitize_plain.log 2>&1; rc=$?
echo "plain rc=$rc"; tail -3 $O/r02s_sanitize_plain.log
[ $rc -ne 0 ] && exit $rc
for tool in memcheck initcheck synccheck; do
  timeout ${SAN_TIMEOUT:-330} compute-sanitizer --tool $tool --print-limit 40 --error-exitcode 9 python tools/sanitize.py > $O/r02s_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|workload done" $O/r02s_sanitize_$tool.log | tail -3
done
timeout ${SAN_TIMEOUT:-330} compute-sanitizer --tool racecheck --print-limit 40 --error-exitcode 9 python tools/sanitize.py cfg2 wide > $O/r02s_sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|workload done" $O/r02s_sanitize_racecheck.log | tail -3
