"""SASS evidence for the claims in DESIGN.md: per kernel family, how many instructions of the kinds the design
leans on the built library contains (cuobjdump -sass of cacaoengine_b200/lib/libcacaoimage_b200.so).

    python profiles/sass_counts.py > profiles/sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cacaoengine_b200", "lib", "libcacaoimage_b200.so")
PATTERNS = [
	("LDG.256", r"\bLDG\.E\S*\.256\b"), ("STG.256", r"\bSTG\.E\S*\.256\b"), ("LDG.128", r"\bLDG\.E\S*\.128\b"), ("STG.128", r"\bSTG\.E\S*\.128\b"),
	("LDG.NA (no L1 allocate)", r"\bLDG\.E\S*\.NA\b"), ("LDS", r"\bLDS\b"), ("FFMA2", r"\bFFMA2\b"), ("FADD2", r"\bFADD2\b"), ("FADD.RZ / FADD2.RZ", r"\bFADD2?\.RZ\b"),
	("FADD.RM", r"\bFADD\.RM\b"), ("F2I", r"\bF2I\b"), ("FRND", r"\bFRND\b"), ("MUFU", r"\bMUFU\b"), ("PRMT", r"\bPRMT\b"), ("IMAD.WIDE", r"\bIMAD\.WIDE\b"),
	("UBLKCP (cp.async.bulk)", r"\bUBLKCP\b"), ("UTMALDG (tensor-map TMA)", r"\bUTMALDG\b"), ("SYNCS (mbarrier)", r"\bSYNCS\b"), ("DMUL", r"\bDMUL\b"), ("HMMA/UTCMMA (tensor cores)", r"\b(HMMA|UTC\w*MMA)\b"),
]


def main():
	txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
	funcs = re.split(r"\n\s*Function : ", txt)[1:]
	names = subprocess.run(["c++filt"], input="\n".join(f.split("\n")[0] for f in funcs), capture_output=True, text=True).stdout.split("\n")
	fam = collections.defaultdict(lambda: {"kernels": 0, "instructions": 0, **{k: 0 for k, _ in PATTERNS}})
	for f, dem in zip(funcs, names):
		m = re.search(r"(\w+)(<|\()", dem.replace("void ", "").replace("cacao::", "").replace("(anonymous namespace)::", ""))
		family = m.group(1) if m else dem
		body = [ln for ln in f.split("\n") if re.match(r"\s*/\*[0-9a-f]{4}\*/", ln)]
		d = fam[family]
		d["kernels"] += 1
		d["instructions"] += len(body)
		joined = "\n".join(body)
		for key, pat in PATTERNS:
			d[key] += len(re.findall(pat, joined))
	print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), instruction counts summed over every instantiation of a kernel family")
	print(f"# {len(funcs)} kernels in the library; zero everywhere: tensor-core instructions (nothing on this path is a contraction) and DMUL outside the gray-weight fallback")
	cols = ["kernels", "instructions"] + [k for k, _ in PATTERNS]
	for family in sorted(fam):
		d = fam[family]
		print(f"\n{family}")
		for k in cols:
			if d[k]:
				print(f"    {k:28s} {d[k]}")


if __name__ == "__main__":
	sys.exit(main())
