"""Derives the per-chromosome LD-block lengths (bp) of the EUR Berisa-Pickrell partition from the
reference's block_data/EUR/chr{1..22}.bed and stores them as package data.  The synthetic
workload generator only needs block LENGTHS (SNP counts per block under a uniform bp density),
not the boundaries themselves.  Run in the build container (needs /root/reference)."""
import json, os, sys
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/block_data/EUR"
out = {}
for c in range(1, 23):
    L = []
    with open(os.path.join(ref, f"chr{c}.bed")) as f:
        for line in f:
            t = line.split()
            if len(t) >= 3 and t[1].isdigit():
                L.append(int(t[2]) - int(t[1]))
    out[str(c)] = L
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dbslmm_b200", "data", "eur_ld_block_lengths.json")
json.dump(out, open(dst, "w"))
print(dst, sum(len(v) for v in out.values()), "blocks")
