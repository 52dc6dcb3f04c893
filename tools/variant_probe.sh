#!/bin/bash
# usage: tools/variant_probe.sh "<nvcc extra flags>[ @ENV=VALUE ...]" ...   (words starting with @ are exported, e.g. @TEEFLOW_SPEC=1.5)
# Rebuilds libteeflow.so per variant, then prints the clip rate (tools/profile_clip.py) and the phase-pure launch
# times (tools/phase_times.py).  The default build is restored at the end.
for spec in "$@"; do
  v=""; envs=()
  for w in $spec; do case "$w" in @*) envs+=("${w#@}");; *) v="$v $w";; esac; done
  for e in "${envs[@]}"; do export "$e"; done
  TEEFLOW_NVCC_EXTRA="$v" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $spec"
  python tools/profile_clip.py 64 3 2>&1 | tail -1 | sed 's/.*pairs\/s/  clip pairs\/s/'
  python tools/phase_times.py 2>&1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('  phase ms:', {k: round(v['ms'], 4) for k, v in d.items() if isinstance(v, dict)})
    print('  launches :', d.get('launch_ms'))
except Exception as e:
    print('  phase probe failed', e)
"
  for e in "${envs[@]}"; do unset "${e%%=*}"; done
done
python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
