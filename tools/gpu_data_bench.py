"""Data-path measurement: batch assembly from the HBM-resident dataset (cdae_gather_images) against the HBM
roofline, next to the reference's per-item path (PIL PNG decode -> ToTensor -> collate, one worker) timed on the host."""
import io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from causaldiffae_b200 import ops
from tests.golden import dataset_fixture as fx

dev = torch.device("cuda:0")
PEAK = 6543.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

res = []
FIRST = "first" in sys.argv[1:]
for name, n, H, W, C, L in ((("pendulum 96x96x4", 8192, 96, 96, 4, 4),) if FIRST else (("pendulum 96x96x4", 8192, 96, 96, 4, 4), ("circuit 128x128x3", 8192, 128, 128, 3, 4),
                            ("morphomnist 28x28x1", 60000, 28, 28, 1, 2))):
    im = torch.randint(0, 256, (n, H, W, C), dtype=torch.uint8, device=dev)
    lab = torch.rand(n, L, device=dev)
    for B in ((4096,) if FIRST else (64, 1024, 4096)):
        g = torch.Generator(device=dev).manual_seed(B)
        idxs = [torch.randint(0, n, (B,), device=dev, generator=g) for _ in range(8)]
        out = torch.empty(B, C, H, W, device=dev); ol = torch.empty(B, L, device=dev)
        k = [0]
        def f():
            ops.gather_images(im, idxs[k[0] % 8], labels=lab, out=out, out_labels=ol); k[0] += 1
        ms = timeit(f)
        byts = B * H * W * C * 5.0
        res.append(dict(dataset=name, batch=B, us=round(ms * 1e3, 2), img_per_s=round(B / ms * 1e3), GBps=round(byts / ms / 1e6, 1),
                        frac_hbm=round(byts / ms / 1e6 / PEAK, 3)))
        print(json.dumps(res[-1]), flush=True)

if FIRST:
    sys.exit(0)
# reference per-item path on the host (its DataLoader runs ONE worker): PNG decode + ToTensor + stack
from PIL import Image
from torchvision import transforms
rng = np.random.RandomState(0)
imgs = fx._smooth(rng, 64, 96, 96, 4)
blobs = []
for i in range(64):
    b = io.BytesIO(); Image.fromarray(imgs[i], mode="RGBA").save(b, format="PNG"); blobs.append(b.getvalue())
tt = transforms.ToTensor()
t0 = time.perf_counter(); reps = 0
while time.perf_counter() - t0 < 5.0:
    batch = torch.stack([tt(Image.open(io.BytesIO(b))) for b in blobs]); reps += 1
dt = time.perf_counter() - t0
print(json.dumps(dict(cpu_reference_path="PIL PNG decode + ToTensor + stack, 1 worker, pendulum 96x96x4", img_per_s=round(64 * reps / dt),
                      cores=1)), flush=True)
