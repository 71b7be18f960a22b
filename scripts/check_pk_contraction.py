"""ptxas contracts a packed multiply whose only use is a packed add into FFMA2 (pk_scalar.cuh).  This script checks the PRODUCT
kernels themselves: for every k_rollout_pk entry of csrc/mbd_b200.cu no `mul.rn.f32x2` of the PTX may disappear from the SASS
(FMUL2 count >= PTX count: ptxas occasionally DUPLICATES a multiply when it rematerialises, which is harmless; a drop is a contraction).
tests/test_pk_host.py runs it; the small probe TU of the same test file is the quick, exact-count version.
    python scripts/check_pk_contraction.py        -> exit code 0 / 1"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SRC = os.path.join(ROOT, "mbd_b200", "csrc", "mbd_b200.cu")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(ROOT, "mbd_b200", "csrc")]


def main():
    with tempfile.TemporaryDirectory() as tmp:
        ptx, cubin = os.path.join(tmp, "k.ptx"), os.path.join(tmp, "k.cubin")
        p1 = subprocess.Popen(["nvcc"] + FLAGS + ["-ptx", SRC, "-o", ptx])
        p2 = subprocess.Popen(["nvcc"] + FLAGS + ["-cubin", SRC, "-o", cubin])
        if p1.wait() != 0 or p2.wait() != 0:
            print("nvcc failed"); return 2
        text = open(ptx).read()
        entries = {}
        for m in re.finditer(r"\.entry\s+(\w+)\s*\(", text):
            name = m.group(1)
            end = text.find("\n}", m.end())
            entries[name] = text[m.end():end].count("mul.rn.f32x2")
        sass = subprocess.run(["cuobjdump", "-sass", cubin], check=True, capture_output=True, text=True).stdout
        bad = 0
        for chunk in re.split(r"\n\s*Function : ", sass)[1:]:
            name = chunk.split("\n", 1)[0].strip()
            if "k_rollout_pk" not in name:
                continue
            n_sass = len(re.findall(r"\bFMUL2\b", chunk))
            n_ptx = entries.get(name, -1)
            ok = n_sass >= n_ptx
            bad += 0 if ok else 1
            print(f"{name[:60]:60s} mul.rn.f32x2 {n_ptx:5d}  FMUL2 {n_sass:5d}  {'ok' if ok else 'CONTRACTED'}")
        return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
