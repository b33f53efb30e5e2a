#!/bin/bash
# Runs every selftest case in its own process (a trapped kernel poisons the CUDA context) under a
# timeout, and prints a PASS/FAIL summary.  Usage: tools/run_selftest.sh [case ...]
cd "$(dirname "$0")/.."
BIN=tools/selftest
mkdir -p gpurun_out
cases="$@"
[ -z "$cases" ] && cases=$($BIN list)
pass=0; fail=0; failed=""
for c in $cases; do
  echo "=== $c"
  timeout 120 $BIN $c
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); else fail=$((fail+1)); failed="$failed $c(rc=$rc)"; fi
done
echo "SELFTEST SUMMARY: pass=$pass fail=$fail failed:$failed"
[ $fail -eq 0 ]
