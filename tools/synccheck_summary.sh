#!/bin/bash
# synccheck over one sanitize case with every error kept, reduced to distinct (message, kernel, source line, barrier address) sites.
# usage: bash tools/synccheck_summary.sh <tag> [case]
tag=$1; c=${2:-T24}
timeout 1200 compute-sanitizer --tool synccheck --print-limit 200000 python tools/sanitize_case.py $c > /tmp/sync_full.log 2>&1
echo "rc=$?" > gpurun_out/${tag}_synccheck_sites_$c.txt
grep -E "checksum|ERROR SUMMARY" /tmp/sync_full.log >> gpurun_out/${tag}_synccheck_sites_$c.txt
grep -A5 "Barrier error" /tmp/sync_full.log | grep -E "Barrier error|located at|Device Frame" | sed -E 's/by thread.*//; s/\+0x[0-9a-f]+//g' | paste - - - - | sort | uniq -c | sort -rn | head -40 >> gpurun_out/${tag}_synccheck_sites_$c.txt
grep "by thread" /tmp/sync_full.log | sed -E 's/.*thread \(([0-9]+),0,0\).*/\1/' | awk '{print int($1/32)}' | sort -n | uniq -c >> gpurun_out/${tag}_synccheck_sites_$c.txt
cat gpurun_out/${tag}_synccheck_sites_$c.txt
