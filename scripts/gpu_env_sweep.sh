#!/bin/bash
# usage: scripts/gpu_env_sweep.sh "<rung_cost args>" "ENV=VAL ENV2=VAL" "..."   (one rung_cost run per environment string)
ARGS=$1; shift
for e in "$@"; do echo "== $e"; env $e timeout 300 python scripts/rung_cost.py $ARGS 2>&1 | tail -6; done
