#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun flags] -- 'command'   — retries while the pod answers busy (exit 3 / transient)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy ($i), sleeping 90 s"; sleep 90
done
exit 3
