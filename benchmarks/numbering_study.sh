#!/bin/bash
# Unstructured-ordering study: C4 with its native (lexicographic) numbering, with vertices and cells
# randomly renumbered, and with the library's Morton + degree-windowed reorder applied to the random numbering.
# One JSON line per case into profiles/r02_numbering_<case>.json
set -e
out=${1:-gpurun_out}
for c in "native none" "random none" "random morton"; do
  set -- $c
  python bench.py --numbering $1 --reorder $2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r02_numbering_$1_$2.json 2> $out/r02_numbering_$1_$2.err || { tail -5 $out/r02_numbering_$1_$2.err; }
  python - "$out/r02_numbering_$1_$2.json" "$1/$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); k = d["kernels"]; c = d["config"]
f = lambda n: "%.3f ms (%.2f)" % (k[n]["ms"], k[n]["frac"]) if "ms" in k.get(n, {}) else "-"
print(sys.argv[2], "| padding %.3f | spmv_kuu %s | smoother %s | fc_kcc_rows %s | fu %s | tile asm %s | value %.2f its_u %.1f setup %.1f s" % (
    k["_pattern"]["sell_padding"], f("spmv_kuu"), f("smoother_step_fp16"), f("fc_kcc_rows"), f("fu_spmv"), f("assembly_full_tile"),
    d["value"], c["krylov_its_u_per_step"], d["setup_s"]))
PY
done
