import re, sys
def load(path):
    fn=None; d={}
    for ln in open(path):
        m=re.search(r'Function : (\S+)', ln)
        if m: fn=m.group(1); d[fn]=[]; continue
        if fn is None: continue
        ln=re.sub(r'/\*[0-9a-f]{4}\*/','',ln)          # instruction addresses
        ln=re.sub(r'/\* 0x[0-9a-f]{16} \*/','',ln)     # encodings (contain relative branch targets -> keep mnemonic text only)
        if ln.strip(): d[fn].append(ln.strip())
    return d
a,b=load(sys.argv[1]),load(sys.argv[2])
print(len(a),"functions vs",len(b))
bad=0
for k in sorted(set(a)|set(b)):
    if k not in a or k not in b: print("ONLY IN ONE:",k); bad+=1; continue
    if a[k]!=b[k]:
        bad+=1
        print("DIFFERS:",k,len(a[k]),len(b[k]))
        for x,y in zip(a[k],b[k]):
            if x!=y: print("   ",x,"|",y); break
print("IDENTICAL PER FUNCTION" if bad==0 else f"{bad} functions differ")
