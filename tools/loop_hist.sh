#!/bin/bash
# SASS opcode histogram of the K loop (largest backward branch) of k_env_step32<false>; usage: tools/loop_hist.sh obj
obj=${1:-b747_rl_ctrl_b200/build/b747_kernels_f32.o}
cuobjdump -sass -fun '_ZN4b74712k_env_step32ILi0ELi4ELb0EEEvNS_6DevCfgENS_4MP32ENS_8StateF32EPKfPfS6_PhS6_' $obj > /tmp/loop.sass
python3 - <<'PY'
import re,collections
L=[l for l in open('/tmp/loop.sass') if re.match(r'\s+/\*[0-9a-f]{4}\*/',l)]
ins=[]
for l in L:
    m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);',l)
    ins.append((int(m.group(1),16),m.group(2).strip()))
loops=[]
for a,t in ins:
    m=re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a: loops.append((tgt,a))
loops.sort(key=lambda l:l[0]-l[1])
# the K loop is the largest loop nested inside the persistent tile loop (or the largest one if there is no tile loop)
best=loops[0]
for l in loops[1:]:
    if l[0]>=loops[0][0] and l[1]<=loops[0][1] and (l[1]-l[0])*3>(loops[0][1]-loops[0][0]):
        best=l; break
print("loop 0x%x..0x%x  %d instructions"%(best[0],best[1],(best[1]-best[0])//16+1))
c=collections.Counter()
for a,t in ins:
    if best[0]<=a<=best[1]:
        t=re.sub(r'^@!?U?P\w+\s+','',t)
        c[t.split()[0].split('.')[0]]+=1
print(", ".join("%s %d"%kv for kv in c.most_common(40)))
PY
