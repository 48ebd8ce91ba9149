"""Prints the key numbers of a bench.py JSON line read from stdin (label = argv[1])."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
c = d["config"]
print(sys.argv[1] if len(sys.argv) > 1 else "bench", "value %.2f steady %.2f cold %.2f e2e %s | krylov %.2f ms asm %.2f ms its_u %.2f its_c %.2f launches/step %.0f setup %.2f s" % (
    d["value"], d["value_steady"], d["value_cold"], ("%.2f" % d["e2e"]["value"]) if d.get("e2e") else "-", c["ms_krylov_per_step"],
    c["ms_assembly_per_step"], c["krylov_its_u_per_step"], c["krylov_its_c_per_step"], c["launches_per_step"], d["setup_s"]))
if d.get("parity_vs_1gpu"):
    print("   parity vs 1 GPU: u %.2e c %.2e" % (d["parity_vs_1gpu"]["rel_l2_u"], d["parity_vs_1gpu"]["rel_l2_c"]))
