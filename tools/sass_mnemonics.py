"""Counts the tcgen05 / TMA / reduction SASS mnemonics per kernel of the given object files (cuobjdump -sass):
    python tools/sass_mnemonics.py pinn_based_online_pde_calculator_b200/csrc/build/tc_128_202.o ...
UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA bulk copy), UBLKPF = bulk L2 prefetch,
UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, REDG = red.global."""
import re, subprocess, sys, collections
pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCMMA|UTCBAR|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|UTMASTG|UTMAPF|REDG|RED|HMMA|SYNCS|UTCATOMSWS|FENCE)\b[\.\w]*')
for obj in sys.argv[1:]:
    out = subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
    fn=None; cnt=collections.OrderedDict()
    for line in out.splitlines():
        m=re.search(r'Function : (\S+)',line)
        if m:
            fn=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip(); cnt[fn]=collections.Counter(); continue
        m=pat.search(line)
        if m and fn: cnt[fn][m.group(0)]+=1
    print('==',obj)
    for fn,c in cnt.items():
        print(' ',fn[:110]); 
        for k,v in sorted(c.items()): print('      %5d  %s'%(v,k))
