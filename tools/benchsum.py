import json,sys
d=json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
print("value %.3e kp/s  ms/step %.3f | e2e %.3e (%.2f ms) | e2e_u8 %.3e (%.2f ms) | tracked %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_u8_host_frames"]["value"], d["e2e_u8_host_frames"]["ms_per_step"], d["tracked_fraction"]))
s=d["step_roofline"]; print("build %.3f ms (%.1f%% hbm)  track %.3f ms | step frac %.3f" % (s["build_ms"], 100*s["pyramid_build_frac"], s["track_ms"], s["frac_of_hbm_peak"]))
print("roofline", d["roofline"]["kernel"], "%.1f GB/s frac %.3f" % (d["roofline"]["achieved"], d["roofline"]["frac"]), "| lk", d["lk_fp32"])
print(" ".join("%s=%.3f" % (k.replace("k_",""), v["ms_per_launch"]) for k,v in sorted(d["kernels"].items())))
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
print("clocks", d["clocks"], "launches", d["gpu_launches"])
