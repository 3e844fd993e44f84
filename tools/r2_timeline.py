"""Timeline of ONE steady-state training step (graph replay + optimizer) from torch.profiler (CUPTI activity records):
   python tools/r2_timeline.py out.csv   -> rows: start_us (relative), dur_us, stream, kernel name; prints a summary:
   wall time of the step, busy time per stream, idle gaps on the union of streams, per-kernel warm durations."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from causaldiffae_b200 import script_util as su, dist_util, logger
from causaldiffae_b200.train_util import TrainLoop
import causaldiffae_b200.nn as cnn

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_timeline.csv"
what = sys.argv[2] if len(sys.argv) > 2 else "train"        # "ddim": two DDIM steps at 512 interventions instead
_lr = int(os.environ.get("LOCAL_RANK", "0"))          # under torchrun: every rank runs, rank 0 reports
torch.cuda.set_device(_lr)
dev = torch.device("cuda", _lr)
dist_util.setup_dist()
logger.configure(dir="/tmp/cdae_prof", format_strs=[])
cnn.RNG_MODE = "device"
torch.manual_seed(0)
model, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **bench.FLAGS}, A=bench.PENDULUM)
bench._dezero(model)
model.to(dev)
if what == "ddim":
    from causaldiffae_b200.sampling import counterfactual
    from torch.profiler import profile, ProfilerActivity
    import json, types
    _, d5 = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **bench.FLAGS, "timestep_respacing": "ddim5"},
                                          A=bench.PENDULUM)
    model.eval()
    xs = torch.rand(512, 3, 64, 64, device=dev)
    counterfactual(model, d5, xs, do_var=0, do_value=0.2)
    counterfactual(model, d5, xs, do_var=0, do_value=0.2)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        counterfactual(model, d5, xs, do_var=0, do_value=0.1)
        torch.cuda.synchronize()
    trace = out.replace(".csv", "_trace.json")
    prof.export_chrome_trace(trace)
    ev = sorted(((float(t["ts"]), float(t["dur"]), t["name"]) for t in json.load(open(trace))["traceEvents"]
                 if t.get("cat") == "kernel" and t.get("ph") == "X"))
    os.remove(trace)
    t0 = ev[0][0]
    wall = max(s + d for s, d, _ in ev) - t0
    busy = sum(d for _, d, _ in ev)
    print(f"encode + 5 DDIM steps at 512 interventions: wall {wall:.0f} us, {len(ev)} kernels, sum of kernel durations {busy:.0f} us")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for s_, d, n in ev:
        k = n.split("(")[0].replace("void ", "").replace("cdae::", "")[:44]
        agg[k][0] += d; agg[k][1] += 1
    for k, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"{d:9.1f} us {c:4d}  {100 * d / busy:5.1f}%  {k}")
    with open(out, "w") as f:
        f.write("start_us,dur_us,name\n")
        for s_, d, n in ev:
            f.write(f"{s_ - t0:.1f},{d:.1f},\"{n[:90]}\"\n")
    sys.exit(0)
B = 64
loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                 log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=4,
                 causal_modeling=True, in_channels=3)
np.random.seed(0)
x, cond = bench.synth_batch(B, 1, device=dev)
for _ in range(8):
    loop.run_step(x, dict(cond))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        loop.run_step(x, dict(cond))
    torch.cuda.synchronize()
import json, types
if int(os.environ.get("RANK", "0")) != 0:          # under torchrun the other ranks only take part in the collectives
    import torch.distributed as _d
    _d.barrier()
    sys.exit(0)
trace = out.replace(".csv", "_trace.json")
prof.export_chrome_trace(trace)
tj = json.load(open(trace))
ev = []
for t in tj["traceEvents"]:
    if t.get("cat") == "kernel" and t.get("ph") == "X":
        e = types.SimpleNamespace(name=t["name"], stream=t.get("args", {}).get("stream", 0),
                                  time_range=types.SimpleNamespace(start=float(t["ts"]), end=float(t["ts"]) + float(t["dur"])))
        ev.append(e)
os.remove(trace)
ev.sort(key=lambda e: e.time_range.start)
# the last step = the events after the last q_sample kernel
starts = [i for i, e in enumerate(ev) if "pack_weights" in e.name or "q_sample" in e.name]
first = [i for i, e in enumerate(ev) if "randn" in e.name]
i0 = max(i for i in first if all(j > i or j < i - 5 for j in first if j != i) or True)
# simpler: split by the adam kernel
adams = [i for i, e in enumerate(ev) if "adam_tick" in e.name]
lo = adams[-2] + 1 if len(adams) >= 2 else 0
hi = adams[-1] + 1
step = ev[lo:hi]
t0 = step[0].time_range.start
rows = [(e.time_range.start - t0, e.time_range.end - e.time_range.start, e.stream, e.name) for e in step]
with open(out, "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for s, d, st, n in rows:
        f.write(f"{s:.1f},{d:.1f},{st},\"{n[:90]}\"\n")
wall = max(s + d for s, d, _, _ in rows)
busy = sum(d for _, d, _, _ in rows)
# union coverage
iv = sorted((s, s + d) for s, d, _, _ in rows)
cov, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        cov += cur_e - cur_s; gaps.append((s - cur_e, cur_e)); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
cov += cur_e - cur_s
print(f"step wall {wall:.0f} us, kernels {len(rows)}, sum of kernel durations {busy:.0f} us, union busy {cov:.0f} us, idle {wall - cov:.0f} us in {len(gaps)} gaps")
per_stream = collections.defaultdict(float)
for s, d, st, n in rows:
    per_stream[st] += d
print("busy per stream (us):", {k: round(v) for k, v in per_stream.items()})
agg = collections.defaultdict(lambda: [0.0, 0])
for s, d, _, n in rows:
    k = n.split("(")[0].replace("void ", "").replace("cdae::", "")[:44]
    agg[k][0] += d; agg[k][1] += 1
for k, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"{d:9.1f} us {c:4d}  {100 * d / busy:5.1f}%  {k}")
gaps.sort(reverse=True)
print("largest gaps (us, at):", [(round(g, 1), round(a)) for g, a in gaps[:12]])
nccl = [(round(s_), round(d_)) for s_, d_, _, n_ in rows if "nccl" in n_.lower()]
if nccl:
    print("NCCL kernels (start_us, dur_us):", nccl)
    tail = [(round(s_), round(d_), n_.split("(")[0][-40:]) for s_, d_, _, n_ in rows if s_ > rows[-1][0] - 1500]
    print("last 1.5 ms:", tail[-40:])
print("gap total by size: >5us", round(sum(g for g, _ in gaps if g > 5)), " 2-5us", round(sum(g for g, _ in gaps if 2 < g <= 5)), " <=2us", round(sum(g for g, _ in gaps if g <= 2)))
import torch.distributed as _d
if _d.is_initialized() and _d.get_world_size() > 1:
    _d.barrier()
