import re,sys,subprocess
obj,fn=sys.argv[1],sys.argv[2]
out=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout.splitlines()
on=False;rows=[]
i=0
while i<len(out):
    l=out[i]
    if 'Function :' in l: on = fn in l
    elif on:
        m=re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/',l)
        if m and i+1<len(out):
            m2=re.match(r'\s*/\* (0x[0-9a-f]+) \*/',out[i+1])
            if m2:
                hi=int(m2.group(1),16)
                stall=(hi>>41)&0xf; yld=(hi>>45)&1; wbar=(hi>>46)&7; rbar=(hi>>49)&7; wait=(hi>>52)&0x3f
                rows.append((m.group(1),m.group(2).strip(),stall,yld,wbar,rbar,wait))
                i+=1
    i+=1
lo=int(sys.argv[3],16); hi_=int(sys.argv[4],16)
for a,ins,stall,yld,wbar,rbar,wait in rows:
    if lo<=int(a,16)<=hi_:
        print(f"{a} st={stall:2d} {'Y' if yld else ' '} w={'-' if wbar==7 else wbar} r={'-' if rbar==7 else rbar} wait={wait:06b}  {ins[:70]}")
