#!/bin/bash
# static SASS instruction count per source line of one kernel: static_lines.sh <kernel substring> [topN]
set -e
rm -rf /tmp/xelf && mkdir /tmp/xelf && cd /tmp/xelf
cuobjdump -xelf all /root/repo/2022_cambroise_interpret_multivae_b200/libmopoe_b200.so >/dev/null 2>&1
nvdisasm --print-line-info mopoe_daa.sm_100a.cubin 2>&1 | awk -v k="$1" '/\.text\./{f=($0 ~ k)} f' > /tmp/k_lines.sass
python3 - "$2" <<'PY'
import re, sys, collections
top = int(sys.argv[1] or 30)
cur=None; cnt=collections.Counter()
for ln in open('/tmp/k_lines.sass'):
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', ln) and cur: cnt[cur]+=1
print("total", sum(cnt.values()))
for k,v in cnt.most_common(top): print("%-24s:%4d %6d" % (k[0],k[1],v))
PY
