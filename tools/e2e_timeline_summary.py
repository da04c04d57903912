"""Summarise gpurun_out/e2e_trace.json (tools/e2e_timeline.py): per-stream busy time and the kernels of one steady step."""
import collections
import json
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/e2e_trace.json"
d = json.load(open(path))
ev = [e for e in d["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
f0 = [e for e in ev if "spmm_rowsplit_kernel<4, 5" in e["name"]]
print("fwd0 start-to-start (us):", [round(b["ts"] - a["ts"]) for a, b in zip(f0, f0[1:])])
rt = collections.Counter(e["name"] for e in d["traceEvents"] if e.get("cat") == "cuda_runtime")
print("cudaMalloc calls in the trace:", rt.get("cudaMalloc", 0))
streams = collections.Counter(e["args"].get("stream") for e in ev)
main = streams.most_common(1)[0][0]
k = int(sys.argv[2]) if len(sys.argv) > 2 else len(f0) - 6
t0, t1 = f0[k]["ts"], f0[k + 1]["ts"]
print(f"\none step while the prefetch worker is busy ({t1 - t0:.0f} us), main stream = s{main}; main-stream events < 6 us omitted")
print("| t (us) | dur (us) | stream | kernel |\n|---:|---:|---|---|")
busy = collections.Counter()
for e in ev:
    if e["ts"] < t0 or e["ts"] >= t1:
        continue
    s = e["args"].get("stream")
    busy[s] += e["dur"]
    if s == main and e["dur"] < 6:
        continue
    name = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
    print(f"| {e['ts'] - t0:.0f} | {e['dur']:.1f} | s{s} | `{name[:70]}` |")
print("\nbusy us per stream in that step:", {f"s{k_}": round(v) for k_, v in busy.items()})
