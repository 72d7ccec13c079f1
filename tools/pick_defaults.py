"""Summarise A/B runs (tools/quick_ab.py jsonl or tools/ab_variants.py json lines): best variant per mesh, speed-up over
the current default, PCG variants, SELL-sigma effect.

    python tools/pick_defaults.py gpurun_out/r2a_quick_ab.jsonl [more files ...]
"""
import json
import sys

DEFAULT = {"C3D4": 5, "C3D10": 1}


def main(paths):
    asm, cg, cg10 = {}, {}, {}
    for path in paths:
        for line in open(path):
            line = line.strip()
            if not line or not line.startswith("{"):
                continue
            d = json.loads(line)
            if d.get("what") == "assembly" and "ms" in d:
                asm.setdefault(d["kind"], {})[d["variant"]] = min(d["ms"][1:])
            elif d.get("what") == "cg" and "ms_per_iter" in d:
                cg[d["variant"]] = min(d["ms_per_iter"][1:])
            elif d.get("what") == "cg_c3d10" and "ms_per_iter" in d:
                cg10[d["sigma"]] = (min(d["ms_per_iter"][1:]), d.get("nnzb"), d.get("nslots"))
            elif "assembly" in d and "kind" in d:                       # ab_variants.py line
                for k, v in d["assembly"].items():
                    if "ms" in v:
                        asm.setdefault(d["kind"], {})[int(k[1:])] = v["ms"]
                for k, v in d.get("cg", {}).items():
                    if "ms_per_iter" in v:
                        cg[k] = v["ms_per_iter"]
    for kind, res in asm.items():
        base = res.get(DEFAULT.get(kind, 1))
        print(f"== assembly {kind}: default variant {DEFAULT.get(kind, 1)} = {base} ms")
        for v, ms in sorted(res.items(), key=lambda kv: kv[1]):
            rel = f"{base / ms:5.2f}x" if base else "  -  "
            print(f"   variant {v:2d}: {ms:8.3f} ms  {rel}")
    if cg:
        base = cg.get("persistent")
        print("== PCG (1 GPU), ms/iteration")
        for k, ms in sorted(cg.items(), key=lambda kv: kv[1]):
            print(f"   {k:26s} {ms:.5f}  {base / ms if base else float('nan'):5.3f}x")
    if cg10:
        print("== PCG on the C3D10 matrix vs SELL sigma (ms/iteration, padding)")
        for sg, (ms, nnzb, nslots) in sorted(cg10.items()):
            pad = f"{100 * (1 - nnzb / nslots):.1f}%" if nnzb and nslots else "?"
            print(f"   sigma {sg:5d}: {ms:.5f}  padding {pad}")


if __name__ == "__main__":
    main(sys.argv[1:])
