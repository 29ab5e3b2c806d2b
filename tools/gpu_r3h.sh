#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 120 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/r3h_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 $O/r3h_memcheck.log
timeout 90 compute-sanitizer --tool synccheck python tools/sanitize_run.py > $O/r3h_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -n 3 $O/r3h_synccheck.log
