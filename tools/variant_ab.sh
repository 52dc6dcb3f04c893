#!/bin/bash
# usage: tools/variant_ab.sh "<nvcc extra flags>[ @ENV=VALUE ...]" ...   -- rebuilds per variant, prints tools/ab_clip.py's line
for spec in "$@"; do
  v=""; envs=()
  for w in $spec; do case "$w" in @*) envs+=("${w#@}");; *) v="$v $w";; esac; done
  for e in "${envs[@]}"; do export "$e"; done
  TEEFLOW_NVCC_EXTRA="$v" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== [$spec] $(python tools/ab_clip.py 2>&1 | tail -1)"
  for e in "${envs[@]}"; do unset "${e%%=*}"; done
done
python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
