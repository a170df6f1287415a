#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python tools/profile_step.py --out $out/r3u_timeline > $out/r3u_timeline.log 2>&1; echo "timeline exit $?"; head -14 $out/r3u_timeline.log | tail -10
