"""dev tool: matrices + reference kernel-file hashes for tools/umma_cli_check.sh.  Runs the unmodified
sequential reference (oracle/_ref/lanczos_modp_seq, built by `make -C oracle`) HERE; the GPU box only
compares hashes.  Output: tools/data/{a,b,c}.mtx and tools/data/cases.txt (git-ignored, travels with gpurun)."""
import hashlib, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blk_lanczos_b200 as B

D = os.path.join(ROOT, "tools", "data")
os.makedirs(D, exist_ok=True)
s = B.synth
s.write_mtx(f"{D}/a.mtx", s.uniform_rows(3000, 2900, 20, seed=3))
s.write_mtx(f"{D}/b.mtx", s.uniform_nnz(2500, 3100, 40000, seed=4, order="col"))
s.write_mtx(f"{D}/c.mtx", s.powerlaw_rows(6000, 6000, mean=12.0, seed=5, with_empty_rows=50))
CASES = [("a", 65537, 16, "--left"), ("a", 2147483647, 16, "--right"), ("b", 2147483647, 12, "--right"),
         ("b", 1073741789, 16, "--left"), ("c", 2147483647, 13, "--left"), ("c", 65537, 16, "--right"),
         ("c", 2147483647, 8, "--right"), ("a", 65537, 32, "--left")]
ref = os.path.join(ROOT, "oracle", "_ref", "lanczos_modp_seq")
with open(f"{D}/cases.txt", "w") as f, tempfile.TemporaryDirectory() as tmp:
    for m, p, n, side in CASES:
        out = os.path.join(tmp, "k.mtx")
        r = subprocess.run([ref, "--matrix", f"{D}/{m}.mtx", "--prime", str(p), "--n", str(n), side, "--output-file", out],
                           capture_output=True, text=True, check=True)
        it = r.stdout.split("after ")[1].split(" iterations")[0]
        f.write(f"{m} {p} {n} {side} {hashlib.sha256(open(out, 'rb').read()).hexdigest()} {it}\n")
        print(m, p, n, side, it, flush=True)
