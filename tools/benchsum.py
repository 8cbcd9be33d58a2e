import json,sys
d=json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
e8=d.get("e2e_u8") or d.get("e2e_u8_host_frames")
print("%s | value %.3e kp/s  ms/step %.3f | e2e %.3e (%.2f ms) | e2e_u8 %.3e (%.2f ms) | tracked %.4f" % (d["config"]["workload"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], e8["value"], e8["ms_per_step"], d["tracked_fraction"]))
s=d["step_roofline"]; print("build %.3f ms (%.1f%% hbm)  track %.3f ms  detect %s | step frac %.3f" % (s["build_ms"], 100*s["pyramid_build_frac"], s["track_ms"], s.get("detect_ms_wall"), s["frac_of_hbm_peak"]))
print("roofline", d["roofline"]["kernel"], "%.1f GB/s frac %.3f" % (d["roofline"]["achieved"], d["roofline"]["frac"]), "| lk", {k: v for k, v in (d.get("lk_issue") or d.get("lk_fp32")).items() if k != "ncu"})
print(" ".join("%s=%.3f" % (k.replace("k_",""), v["ms_per_launch"]) for k,v in sorted(d["kernels"].items())))
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
print("clocks", d["clocks"], "launches", d["gpu_launches"], "gather_us", d.get("gather_us"))
