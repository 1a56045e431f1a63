#!/usr/bin/env python3
"""Per-kernel SASS comparison of two builds of libdspmap_b200.so (addresses / immediates normalised):

    python profiles/sass_diff.py OLD.so NEW.so

Used at the end of round 1 to show that the kernels of the default path are, instruction for instruction, the ones of the
last build that ran the GPU suite and the A/B run (commit 7a2fee5): 42 kernels identical; k_cz_chain differs by the position
of one instruction; k_pair_prep gained the branch for the column-major sizing; everything else that differs is reachable
only through an experiment switch."""
import subprocess,re,sys,collections
def kernels(so):
    out=subprocess.run(["cuobjdump","-sass",so],capture_output=True,text=True).stdout
    ks=collections.OrderedDict(); cur=None
    for line in out.splitlines():
        m=re.search(r"Function : (\S+)",line)
        if m: cur=m.group(1); ks[cur]=[]; continue
        m=re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);",line)
        if m and cur: 
            ins=m.group(1).strip()
            ins=re.sub(r"0x[0-9a-f]+","0x",ins)   # addresses / immediates that shift with layout
            ks[cur].append(ins)
    return ks
a=kernels(sys.argv[1]); b=kernels(sys.argv[2])
def norm(n): 
    n=re.sub(r"k_weight2_tILb0EEv","k_weight2",n); n=re.sub(r"k_weight2w_tILb0EEv","k_weight2w",n)
    n=re.sub(r"ILi256ELi8192ELi128ELb0EE","ILi256ELi8192ELi128EE",n); n=re.sub(r"ILi128ELi4096ELi128ELb0EE","ILi128ELi4096ELi128EE",n)
    return n
bn={norm(k):v for k,v in b.items()}
same=diff=0
for k,v in a.items():
    kk=norm(k)
    # old names of the weight kernels
    kk2=kk.replace("_Z9k_weight28MapConst","_Z11k_weight28MapConst").replace("_Z10k_weight2w8MapConst","_Z12k_weight2w8MapConst")
    w=bn.get(kk) or bn.get(kk2)
    if w is None:
        cand=[x for x in bn if re.sub(r"^_Z\d+","",x)[:12]==re.sub(r"^_Z\d+","",kk)[:12]]
        print("not found in new:",k,cand[:2]); continue
    if v==w: same+=1
    else:
        diff+=1
        import difflib
        d=[l for l in difflib.unified_diff(v,w,lineterm="",n=0) if not l.startswith(("---","+++","@@"))]
        print("DIFF %-70s old %d new %d instrs, %d changed lines"%(k[:70],len(v),len(w),len(d)))
print("identical kernels:",same,"different:",diff)
