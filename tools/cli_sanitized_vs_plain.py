"""Dev tool (tools/sanitize.sh): halLiftover / hal2maf / halAlignmentDepth built with sanitizers (argv[1] = their directory) against
the plain emulator builds in tests/simt on random inputs over argv[2]: same output, same messages, no sanitizer report."""
import os, random, subprocess, sys, tempfile
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT+"/oracle"); import pyoracle
A=sys.argv[1]; N=ROOT+"/tests/simt"
hal=sys.argv[2]; o=pyoracle.Oracle(hal); rng=random.Random(3); d=tempfile.mkdtemp(); bad=tot=0
def run(b, args):
    r=subprocess.run([b]+args, capture_output=True, text=True)
    return r
def chk(name, args, outs):
    global bad, tot
    tot+=1
    res=[]
    for base in (A,N):
        for f in outs:
            if os.path.exists(f): os.remove(f)
        r=run(f"{base}/{name}", args)
        txt="".join(open(f).read() if os.path.exists(f) else "<none>" for f in outs)
        err="\n".join(l for l in r.stderr.splitlines() if not l.startswith("[halgpu"))
        res.append((r.returncode, txt, r.stdout, err))
    if "Sanitizer" in res[0][3] or "runtime error" in res[0][3] or res[0]!=res[1]:
        bad+=1; print("BAD", name, args, res[0][0], res[1][0], res[0][3][:1500])
for src in o.genomes:
    for tgt in o.genomes:
        if src==tgt: continue
        seqs=o.sequences(o.genome_id(src))
        bed=os.path.join(d,"i.bed"); out=os.path.join(d,"o.bed")
        with open(bed,"w") as f:
            for i in range(150):
                nm,st,ln=rng.choice(seqs)
                if ln<2: continue
                a=rng.randrange(ln); b=min(ln,a+1+rng.randrange(1,400))
                w=rng.choice([3,4,6])
                row=[nm,str(a),str(b)]
                if w>=4: row.append(f"n{i}")
                if w>=6: row+= [str(rng.randrange(1000)), rng.choice("+-")]
                f.write("\t".join(row)+"\n")
        for extra in ([], ["--noDupes"], ["--outPSL"], ["--columnLiftover"] if False else ["--outPSLWithName"]):
            env_t=str(rng.choice([0,1,3]))
            os.environ["HALGPU_TEXT_THREADS"]=env_t
            chk("halLiftover_emul", extra+[hal,src,bed,tgt,out], [out])
    # maf + depth with this genome as reference
    seqs=o.sequences(o.genome_id(src))
    nm,st,ln=rng.choice(seqs)
    maf=os.path.join(d,"o.maf")
    for extra in ([], ["--noDupes"], ["--unique"], ["--noAncestors","--onlyOrthologs"], ["--maxBlockLen","7"]):
        chk("hal2maf_emul", [hal,maf,"--refGenome",src,"--refSequence",nm]+extra, [maf])
    wig=os.path.join(d,"o.wig")
    for extra in ([], ["--noAncestors"], ["--step","3"], ["--countDupes"]):
        chk("halAlignmentDepth_emul", [hal,src,"--outWiggle",wig]+extra, [wig])
print(tot,"cases",bad,"bad")
