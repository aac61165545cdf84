import collections, csv, os, re, subprocess, sys, tempfile
ROOT="/root/repo"
rep, prefix = sys.argv[1], sys.argv[2]
steps = float(sys.argv[3])
tmp = tempfile.mkdtemp()
so = os.path.join(ROOT, "mchap_b200", "_lib", "libmchap_b200.so")
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
st = [i for i, l in enumerate(dis) if l.startswith(".text." + prefix)][0]
en = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and i > st]
en = en[0] if en else len(dis)
seq=[]; cur="kernel"
for l in dis[st:en]:
    m = re.match(r'^(\$?[_A-Za-z][\w$]*):', l)
    if m and not m.group(1).startswith('.L'):
        cur = m.group(1)
    elif re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s", l): seq.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]; ci = hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > ci]
assert len(data)==len(seq),(len(data),len(seq))
ex=[int(r[ci])/steps for r in data]
tot=collections.Counter(); hot=collections.Counter(); dyn=collections.Counter(); hot2=collections.Counter()
for k,e in zip(seq,ex):
    tot[k]+=1; dyn[k]+=e
    if e>=0.05: hot[k]+=1
    if e>=0.5: hot2[k]+=1
for k in tot:
    nm=k
    m=re.search(r'\$(_ZN4mchb\w+)',k)
    print("%5d static %5d hot>=.05 %5d hot>=.5 %8.1f dyn  %s"%(tot[k],hot[k],hot2[k],dyn[k],subprocess.run(['c++filt',k.strip('$').split('$')[-1]],capture_output=True,text=True).stdout.strip()[:100]))
