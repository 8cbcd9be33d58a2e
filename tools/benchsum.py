"""One line per bench JSON file: device / e2e numbers (development helper)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        e, u = d["e2e"], d["e2e_u8"]
        print(f"{f}: N={d['n_gpus']} dev {d['ms_per_step']:.3f} ms {d['value']/1e6:.1f} M | e2e f64 {e['ms_per_step']:.3f} ms {e['value']/1e6:.1f} M "
              f"(sync {e.get('ms_per_step_one_synchronous_call', 0):.3f}) thr {e.get('host_threads')} rates {e.get('upload_engines_rank0')} | "
              f"u8 {u['ms_per_step']:.3f} ms {u['value']/1e6:.1f} M | gather {d.get('gather_us')}")
    except Exception as ex:
        print(f"{f}: ERR {ex}")
