#!/usr/bin/env python
"""Randomised differential run of the bitstream kernel on the CPU SIMT emulator (tests/sim/) against
the oracle: seven flat patterns, haystacks built from small alphabets with long single-byte runs
(exact-carry tiles, tiles without sync bytes, overlapping candidates) at sizes around the tile and
chunk edges; the default kernel and the -DCGX_PAIR=1 variant.  TEST TOOLING, not part of the suite
(run one process at a time: the emulator libraries are built on first use).
    python tools/fuzz_sim.py [seed] [seconds]"""
import random, sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import sim_lib
from oracle_lib import Oracle
pats=[r"\d+\.\d+\.\d+\.\d+", r"\w+@\w+\.\w+", r"[a-z]+=\d+", r"a+ba", r"x\d*y?z", r"[a-c]+[x-z]?", r"fo+\d+b", r"\d+", r"[a-z]+/\d+"]
alph={pats[0]:[b"0123456789. ", b"01.", b"9.\n", b"12345.x/"], pats[1]:[b"ab@. ", b"a@.\n", b"abc_@.-"], pats[2]:[b"ab=12 ", b"a=1", b"az=09\n;"],
      pats[3]:[b"ab ", b"ab"], pats[4]:[b"x1yz ", b"xyz09"], pats[5]:[b"abcxyz ", b"acxz"], pats[6]:[b"fo1b ", b"fo0b9"],
      pats[7]:[b"01 ", b"0123456789abc \n", b"7"], pats[8]:[b"ab/12 ", b"a/1", b"az/09\n;"]}
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
t0=time.time(); n=0
orc={p:Oracle(p) for p in pats}
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 120:
    p=rng.choice(pats)
    a=rng.choice(alph[p])
    # sizes around the tile (4 KB), chunk (8 KB) and gang edges of the current kernel shape
    size=rng.choice([30, 70, 130, 2048, 4090, 4100, 8180, 8200, 12300, 16380, 24580, 40000, 66000])+rng.randrange(0,40)
    # mix: long runs to trigger exact path and open tiles
    parts=[]; tot=0
    while tot<size:
        if rng.random()<0.15:
            k=rng.randrange(1,300); ch=bytes([rng.choice(a)])*k
        else:
            k=rng.randrange(1,20); ch=bytes(rng.choice(a) for _ in range(k))
        parts.append(ch); tot+=len(ch)
    hay=np.frombuffer(b"".join(parts)[:size],dtype=np.uint8)
    want=orc[p].find_all(hay)
    for tiles,defs,tag in ((1,"",""),(1,"-DCGX_PAIR=1","pair")):
        tot_,_,pairs=sim_lib.scan(p,hay,grid=rng.choice([1,2,3]),jit=True,tiles=tiles,defs=defs,tag=tag)
        if tot_!=len(want) or not np.array_equal(pairs,want):
            print("MISMATCH",p,tiles,tag,size,bytes(hay[:200])); open('/tmp/fuzz_fail.bin','wb').write(bytes(hay)); sys.exit(1)
    n+=1
print("ok",n,"cases")
