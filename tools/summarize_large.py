import json, sys
for l in open(sys.argv[1]):
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    print(d["m"], "nstates", d["nstates"], "rhs_ms", round(d["rhs_ms"], 4), "jvp_ms", round(d["jvp_ms"], 4),
          "stream GB/s rhs/jvp", round(d.get("rhs_stream_gbs", 0)), round(d.get("jvp_stream_gbs", 0)),
          "jvp_fd", "%.1e" % d["jvp_vs_central_difference"], "build_s", round(d["build_seconds"], 1))
