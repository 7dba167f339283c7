"""Per-kernel-family milliseconds of one reverse step from a bench.py JSON line (roofline.by_family + roofline_hbm)."""
import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"== {path}: {d['ms_per_step']:.2f} ms/step, {d['value']:.2f} samples/s, clocks {d['clocks']['sm_mhz']} MHz")
    r = d["roofline"]
    print(f"   conv {r['conv_ms_per_step']:.2f} ms ({r['achieved']:.0f} TF/s, frac {r['frac']:.3f}), non-conv {r['non_conv_ms_per_step']:.2f} ms, whole-step frac {r['whole_step_frac']:.3f}")
    for k, v in r["by_family"].items():
        print(f"   {k:24s} n={v['launches']:3d} {v['ms']:7.3f} ms {v['tflops']:7.0f} TF/s")
    for h in d["roofline_hbm"]:
        print(f"   {h['kernel']:24s} n={h['launches_per_step']:3d} {h['ms_per_step']:7.3f} ms {h['achieved']:7.0f} GB/s frac {h['frac']:.2f}")
