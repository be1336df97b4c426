#!/bin/bash
# tools/ab1.sh "<kbench args>" "<env assignments>" ... : one configuration, several library variants
args="$1"; shift
for envs in "$@"; do
  echo "## $envs :: $(env $envs python tools/kbench.py $args --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-140)"
done
