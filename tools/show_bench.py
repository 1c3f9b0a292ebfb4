import json, sys
for f in sys.argv[1:]:
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j.get("roofline") or {}
    ss=j.get("steady_state") or {}
    print(f"{f}: steady={ss.get('value',0)/1e9:.3f} G ({ss.get('ms_per_generation',0):.4f} ms/gen) N={j['n_gpus']} value={j['value']/1e9:.3f} G  ms/gen={j['ms_per_step']/10:.4f}  launches={j['gpu_launches']}  e2e={(j.get('e2e') or {}).get('value', 0)/1e9:.3f} G")
    if r:
        print("   ms/launch", {k: round(v, 4) for k, v in r["kernel_ms_per_launch"].items()})
        print("   launches/gen", r.get("launches_per_generation"))
        if "exchange" in r:
            x = r["exchange"]; print(f"   exchange: {x['achieved']:.0f} GB/s = {x['frac']:.2f} of 770; pushed/gen {x['bytes_pushed_per_generation']/1e6:.1f} MB; send ms/gen {x['ms_per_generation_in_send_kernels']:.4f}")
    if "sharded_check" in j: print("   sharded_check", j["sharded_check"]["bit_identical"], j["sharded_check"]["per_mode"])
