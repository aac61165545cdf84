import collections, csv, os, re, subprocess, sys, tempfile
ROOT="/root/repo"
rep, prefix = sys.argv[1], sys.argv[2]
steps = float(sys.argv[3]); thr=float(sys.argv[4]); top=int(sys.argv[5]); thr2=float(sys.argv[6]) if len(sys.argv)>6 else 1e9
tmp = tempfile.mkdtemp()
so = os.path.join(ROOT, "mchap_b200", "_lib", "libmchap_b200.so")
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
st = [i for i, l in enumerate(dis) if l.startswith(".text." + prefix)][0]
en = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and i > st]
en = en[0] if en else len(dis)
seq, cur = [], ("?", 0)
for l in dis[st:en]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s", l): seq.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]; ci = hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > ci]
assert len(data)==len(seq)
ex=[int(r[ci])/steps for r in data]
b=collections.Counter(); d=collections.Counter()
for k,e in zip(seq,ex):
    if e>=thr and e<thr2: b[k]+=1; d[k]+=e
files={}
for k,v in b.most_common(top):
    f,ln=k
    if f not in files:
        p=os.path.join(ROOT,"mchap_b200","csrc",f); files[f]=open(p).read().split("\n") if os.path.exists(p) else None
    t=files[f]; txt=t[ln-1].strip()[:90] if t and ln-1<len(t) else ""
    print("%4d static %7.1f dyn  %s:%d  %s"%(v,d[k],f,ln,txt))
