"""Filter for dist_check.py output: one compact line per rank-0 record (and any record with an error)."""
import json, sys
for line in sys.stdin.read().replace("}{", "}\n{").splitlines():
    if not line.startswith("{"):
        if "rror" in line or "Traceback" in line: print(line[:300])
        continue
    try: r = json.loads(line)
    except Exception: continue
    if r.get("rank") == 0 or "parity_error" in r or r.get("rel_l2_vs_single_gpu_plan", 0) > 1e-12:
        print({k: r[k] for k in ("driver", "log_n", "world", "rank", "dir", "log_m", "plan_s", "rel_l2_vs_single_gpu_plan", "parity_error", "ms", "gflops", "strict_GBps_aggregate") if k in r})
