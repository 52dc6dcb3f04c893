#!/bin/bash
# usage: tools/variant_clip.sh "<nvcc extra flags>[ @ENV=VALUE ...]" ...   (words starting with @ are exported)
# Rebuilds libteeflow.so per variant and prints the clip rate of tools/profile_clip.py (dataflow kernel).
for spec in "$@"; do
  v=""; envs=()
  for w in $spec; do case "$w" in @*) envs+=("${w#@}");; *) v="$v $w";; esac; done
  for e in "${envs[@]}"; do export "$e"; done
  TEEFLOW_NVCC_EXTRA="$v" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $spec: $(python tools/profile_clip.py 64 4 2>&1 | grep 'pairs/s' | tail -2 | sed 's/.*pairs\/s//' | tr '\n' ' ')"
  for e in "${envs[@]}"; do unset "${e%%=*}"; done
done
python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
