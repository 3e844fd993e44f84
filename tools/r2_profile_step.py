"""One steady-state training step (and one DDIM step at 512 interventions) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/r2_profile_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from causaldiffae_b200 import script_util as su, dist_util, logger
from causaldiffae_b200.train_util import TrainLoop
import causaldiffae_b200.nn as cnn

what = sys.argv[1] if len(sys.argv) > 1 else "train"
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist_util.setup_dist()
logger.configure(dir="/tmp/cdae_prof", format_strs=[])
cnn.RNG_MODE = "device"
torch.manual_seed(0)
model, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **bench.FLAGS}, A=bench.PENDULUM)
bench._dezero(model)
model.to(dev)
if what == "train":
    B = 64
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=4,
                     causal_modeling=True, in_channels=3)
    np.random.seed(0)
    x, cond = bench.synth_batch(B, 1, device=dev)
    for _ in range(5):
        loop.run_step(x, dict(cond))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loop.run_step(x, dict(cond))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    from causaldiffae_b200.sampling import counterfactual
    _, d5 = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **bench.FLAGS, "timestep_respacing": "ddim5"},
                                          A=bench.PENDULUM)
    model.eval()
    x = torch.rand(512, 3, 64, 64, device=dev)
    counterfactual(model, d5, x, do_var=0, do_value=0.2)
    torch.cuda.synchronize()
    _, d1 = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **bench.FLAGS, "timestep_respacing": "ddim2"},
                                          A=bench.PENDULUM)
    counterfactual(model, d1, x, do_var=0, do_value=0.2)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    counterfactual(model, d1, x, do_var=0, do_value=0.1)      # 2 DDIM steps: graph replays
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
