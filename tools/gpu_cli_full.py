"""manual: the whole CLI on a chr19-sized BAM + FASTA against the unmodified reference CLI (tables must be identical)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rsicnv_b200 import synth
from bind import REF_BAMTOOL, REF_BIN
L = int(sys.argv[1]) if len(sys.argv) > 1 else synth.CHR19_LEN
d = "/tmp/clifull"; os.makedirs(d, exist_ok=True)
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
synth.write_fasta(d + "/t.fa", "19", fa)
t = time.time(); synth.write_bam(d + "/t.bam", [("19", L)], {0: reads}, level=1, random_seq=7, threads=32); print("write_bam %.1f s, %d bytes" % (time.time() - t, os.path.getsize(d + "/t.bam")), flush=True)
cli = os.path.join(ROOT, "rsicnv_b200", "bin", "rsicnv")
common = ["rsi", "-b", d + "/t.bam", "-f", d + "/t.fa", "-q", "0", "-Q", "10", "-np"]
for extra in ([], [], [], ["-hostdecode"]):
    t = time.time()
    r = subprocess.run([cli] + common + ["-o", d + "/ours.txt"] + extra, capture_output=True, text=True, env=dict(os.environ, RSICNV_TIMING="1"))
    print("ours %s: %.3f s rc=%d" % (" ".join(extra), time.time() - t, r.returncode), [ln for ln in r.stderr.splitlines() if ln.startswith("#timing")], flush=True)
subprocess.run([REF_BAMTOOL, "index", d + "/t.bam"], check=True)
t = time.time(); subprocess.run([REF_BIN] + common + ["-o", d + "/ref.txt"], check=True, capture_output=True); tr = time.time() - t
print("reference CLI: %.2f s" % tr)
tab = lambda p: [ln for ln in open(p).read().splitlines() if not ln.startswith("#input")]
print("identical tables:", tab(d + "/ours.txt") == tab(d + "/ref.txt"), len(tab(d + "/ref.txt")))
