#!/bin/bash
# usage: gpuretry.sh <timeout> <script under profiles/scratch/r02/>   -- retries while the pod answers "transient"/busy
cd /root/repo
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient\|rc=3\|busy"; then
    if echo "$out" | grep -q "status=ok\|status=fail"; then echo "$out" | tail -80; exit $rc; fi
    sleep 90; continue
  fi
  echo "$out" | tail -80
  exit $rc
done
echo "gave up after 30 tries"
