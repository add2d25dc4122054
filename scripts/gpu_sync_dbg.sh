#!/bin/bash
for v in 4 5; do
  compute-sanitizer --tool synccheck --print-limit 1 scripts/abl/sync_probe $v 2>&1 | grep "variant\|Barrier\|by thread\|SUMMARY" | head -5
done
