#!/usr/bin/env bash
cd "$(dirname "$0")/.."
./tools/micro/mbar_racecheck
timeout 120 compute-sanitizer --tool racecheck ./tools/micro/mbar_racecheck > gpurun_out/c22_mbar_racecheck.log 2>&1; tail -12 gpurun_out/c22_mbar_racecheck.log
