#!/bin/bash
# usage: tools/variant_times.sh "<nvcc extra flags>" ...   -- rebuilds libteeflow.so per variant and times one clip
for v in "$@"; do
  TEEFLOW_NVCC_EXTRA="$v" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $v: $(python tools/profile_clip.py 64 3 2>&1 | tail -1 | sed 's/.*pairs\/s/pairs\/s/')"
done
python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
