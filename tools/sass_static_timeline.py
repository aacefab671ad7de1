"""Static issue timeline of a kernel's main loop from its SASS control codes (no GPU needed).

    python tools/sass_static_timeline.py <cubin> <kernel-name-substring>

Every sm_100 instruction carries a stall count (bits 41-44 of the second 64-bit word): the cycles before the same warp may
issue its next instruction.  Their sum over the biggest backward-branch loop is the time ONE warp needs per iteration if no
scoreboard ever makes it wait; 2 x (DFMA + DMUL + DADD) is the time the half-rate FP64 pipe of its scheduler is busy for it.
The ratio is the FP64-pipe share a single warp can claim; profiles/r02_notes.md (section 5) compares it with the measured
pipe utilisation of every step kernel."""
import sys,re,subprocess,collections
cubin=sys.argv[1]; fn_pat=sys.argv[2]
out=subprocess.run(["cuobjdump","-sass",cubin],capture_output=True,text=True).stdout
fn=None; ins=[]; cur=None
for l in out.splitlines():
    m=re.search(r'Function : (\S+)',l)
    if m: fn=m.group(1); continue
    if not (fn and fn_pat in fn): continue
    m=re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/',l)
    if m:
        cur=[int(m.group(1),16),m.group(2),int(m.group(3),16),None]; ins.append(cur); continue
    m=re.search(r'^\s*/\* (0x[0-9a-f]+) \*/',l)
    if m and cur is not None and cur[3] is None: cur[3]=int(m.group(1),16)
# find the biggest backward-branch loop
best=None
for i,(pc,t,w0,w1) in enumerate(ins):
    m=re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<pc and (best is None or pc-tgt>best[1]-best[0]): best=(tgt,pc)
lo=[i for i,x in enumerate(ins) if x[0]==best[0]][0]; hi=[i for i,x in enumerate(ins) if x[0]==best[1]][0]
tot=0; byop=collections.Counter(); cnt=collections.Counter()
for pc,t,w0,w1 in ins[lo:hi+1]:
    st=(w1>>41)&0xf
    op=re.sub(r'^@!?U?P\d+\s+','',t).split()[0].split('.')[0]
    tot+=st; byop[op]+=st; cnt[op]+=1
f64=cnt['DFMA']+cnt['DMUL']+cnt['DADD']
print(fn_pat,"loop instr",hi-lo+1,"sum stall",tot,"FP64",f64,"pipe-time/static = %.2f"%(2*f64/tot))
for op,c in byop.most_common(8): print("  ",op,cnt[op],"stall sum",c,"avg %.2f"%(c/cnt[op]))
