"""Minimal key-value logger with the call surface the hot path uses (ref improved_diffusion/logger.py:212-241,442-472):
log, logkv (last value), logkv_mean (running mean), dumpkvs, getkvs, get_dir, configure.  Values may be CUDA scalars:
they are only synchronised when dumped (the reference calls .item() per value, ~900 syncs per step)."""
import os
import sys
import tempfile
import time
from collections import defaultdict

_STATE = {"dir": None, "kv": {}, "mean": defaultdict(lambda: [0.0, 0]), "vec": defaultdict(lambda: [0.0, 0.0]),
          "formats": ["stdout", "log", "csv"], "csv_keys": None, "hooks": {}}


def configure(dir=None, format_strs=None, comm=None, log_suffix=""):
    if dir is None:
        dir = os.getenv("OPENAI_LOGDIR") or os.path.join(tempfile.gettempdir(), time.strftime("cdae-%Y-%m-%d-%H-%M-%S"))
    os.makedirs(os.path.expanduser(dir), exist_ok=True)
    _STATE["dir"] = os.path.expanduser(dir)
    if format_strs is not None:
        _STATE["formats"] = list(format_strs)
    _STATE["kv"].clear(); _STATE["mean"].clear(); _STATE["vec"].clear()
    _STATE["csv_keys"] = None


def get_dir():
    if _STATE["dir"] is None:
        configure()
    return _STATE["dir"]


def _is_rank0():
    return int(os.environ.get("RANK", "0")) == 0


def log(*args):
    if _is_rank0():
        print(*args, flush=True)


def warn(*args):
    log("WARN:", *args)


def logkv(key, val):
    _STATE["kv"][key] = val


def logkv_mean(key, val):
    acc = _STATE["mean"][key]
    acc[0] = acc[0] + val
    acc[1] += 1


def logkv_mean_n(prefix, sums, counts):
    """vector running mean: accumulates per-bucket sums and counts (device tensors); dumped as {prefix}{i}"""
    acc = _STATE["vec"][prefix]
    acc[0] = acc[0] + sums
    acc[1] = acc[1] + counts


def set_dump_hook(name, fn):
    """fn() -> {key: float}: values a producer accumulates on the device and only reads when the log is dumped"""
    _STATE["hooks"][name] = fn


def _to_float(v):
    return float(v.item()) if hasattr(v, "item") else float(v)


def getkvs():
    out = {k: _to_float(v) for k, v in _STATE["kv"].items()}
    out.update({k: _to_float(s) / max(n, 1) for k, (s, n) in _STATE["mean"].items()})
    for prefix, (s, c) in _STATE["vec"].items():
        sl = s.tolist() if hasattr(s, "tolist") else list(s)
        cl = c.tolist() if hasattr(c, "tolist") else list(c)
        for i, (si, ci) in enumerate(zip(sl, cl)):
            if ci > 0:
                out[f"{prefix}{i}"] = si / ci
    for fn in list(_STATE["hooks"].values()):
        out.update(fn())
    return out


def dumpkvs():
    d = getkvs()
    _STATE["kv"].clear(); _STATE["mean"].clear(); _STATE["vec"].clear()
    if not _is_rank0() or not d:
        return d
    if "stdout" in _STATE["formats"]:
        w = max(len(k) for k in d)
        print("-" * (w + 17))
        for k in sorted(d):
            print(f"| {k:<{w}} | {d[k]:<10.5g} |")
        print("-" * (w + 17), flush=True)
    if "csv" in _STATE["formats"] and _STATE["dir"]:
        path = os.path.join(_STATE["dir"], "progress.csv")
        keys = sorted(d)
        new = not os.path.exists(path) or _STATE["csv_keys"] != keys
        with open(path, "a") as f:
            if new:
                f.write(",".join(keys) + "\n"); _STATE["csv_keys"] = keys
            f.write(",".join(repr(d[k]) for k in keys) + "\n")
    if "log" in _STATE["formats"] and _STATE["dir"]:
        with open(os.path.join(_STATE["dir"], "log.txt"), "a") as f:
            f.write(" ".join(f"{k}={d[k]:.6g}" for k in sorted(d)) + "\n")
    return d
