"""DDIM decode profile: 3 steady-state DDIM steps at batch 128 between cudaProfilerStart/Stop (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import FLAGS, PENDULUM
from causaldiffae_b200 import script_util as su, dist_util
from causaldiffae_b200.sampling import counterfactual
import causaldiffae_b200.nn as cnn
torch.cuda.set_device(0)
dist_util.setup_dist()
cnn.RNG_MODE = "device"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model, _ = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **FLAGS}, A=PENDULUM)
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for n, p in model.named_parameters():
        if float(p.abs().sum()) == 0.0 and p.dim() > 1:
            p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)
model.cuda().eval()
_, d3 = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **FLAGS, "timestep_respacing": "ddim3"}, A=PENDULUM)
x = torch.rand(B, 3, 64, 64, generator=g).cuda()
for _ in range(2):
    counterfactual(model, d3, x, do_var=0, do_value=0.2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
counterfactual(model, d3, x, do_var=0, do_value=-0.3)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"3 DDIM steps at batch {B}: {e0.elapsed_time(e1):.2f} ms ({e0.elapsed_time(e1)/3:.2f} ms/step)")
